#!/usr/bin/env python
"""Pick the metrics that matter out of an `ncu -i X.ncu-rep --page raw --csv` dump (one kernel per row).
usage: tools/ncu_extract.py raw.csv [row]"""
import csv
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "smsp__inst_executed.sum"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    vals = rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
    seen = set()
    for k in KEYS:
        for i, h in enumerate(hdr):
            if (h == k or h.endswith("." + k)) and h not in seen:
                seen.add(h)
                print(f"{h}: {vals[i]} {units[i]}")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_warp_active.pct"):
            try:
                if float(vals[i]) > 2.0:
                    print(f"{h}: {vals[i]} {units[i]}")
            except ValueError:
                pass


if __name__ == "__main__":
    main()
