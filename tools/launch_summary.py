#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of device time per
kernel.  usage: tools/launch_summary.py gpurun_out/launches.csv [--seq N]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    nseq = int(sys.argv[sys.argv.index("--seq") + 1]) if "--seq" in sys.argv else 0
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = row["Kernel Name"]
        m = re.match(r"(?:void )?(?:nnmpc::)?([\w:]+)(<.*)?\(", name)
        short = name[:110] if not m else m.group(1) + (re.sub(r"nnmpc::", "", m.group(2))[:80] if m.group(2) else "")
        agg[short][0] += 1
        agg[short][1] += v
        seq.append((short, v, row["Grid Size"]))
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(seq)} launches, {tot / 1e3:.2f} ms device time (cold-cache, serialised under ncu)")
    print("# share   launches   avg_us   kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / tot * 100:6.2f}%  {v[0]:6d}  {v[1] / v[0]:9.1f}  {k}")
    for s in seq[:nseq]:
        print(s)


if __name__ == "__main__":
    main()
