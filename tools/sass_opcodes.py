#!/usr/bin/env python
"""Opcode histogram of the built library, per kernel family: the SASS evidence that the hot path is tcgen05 + TMEM +
TMA (UTC*MMA, LDTM, UTMALDG, UTCBAR) and where the legacy tensor path (DMMA) is still used.

    python tools/sass_opcodes.py > profiles/sass_opcodes.txt        (needs cuobjdump; no GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "industrial_nnmpc_2021_b200", "csrc", "libnnmpc.so")
WATCH = ("UTCHMMA", "UTCIMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP", "DMMA", "HMMA",
         "IMMA", "SYNCS", "DFMA", "DSETP", "LDGSTS")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur)[:150]
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: opcode counts (static instructions) of the opcodes that matter")
    print("# whole library: " + ", ".join(f"{k} {total[k]}" for k in WATCH if total[k]))
    print()
    for name, c in per.items():
        hits = [(k, c[k]) for k in WATCH[:13] if c[k]]
        if any(k.startswith(("UTC", "UTMA", "LDTM", "DMMA")) for k, _ in hits):
            print(f"{name}\n    " + ", ".join(f"{k} {v}" for k, v in hits) + f"   (instructions: {sum(c.values())})")


if __name__ == "__main__":
    sys.exit(main())
