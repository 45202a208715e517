#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) as a markdown table: one row per profiled launch.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.md
"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_static", "smem"),
    ("launch__waves_per_multiprocessor", "waves"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [(hdr.index(m), lab) for m, lab in WANT if m in hdr]
    print("| kernel | " + " | ".join(f"{lab} ({units[i]})" if units[i] else lab for i, lab in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        print(f"| {name} | " + " | ".join(r[i] for i, _ in cols) + " |")


if __name__ == "__main__":
    main()
