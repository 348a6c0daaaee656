#!/usr/bin/env python3
"""Summarise an ncu report (one kernel launch) into the handful of metrics the roofline uses.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>_summary.txt
"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+)"
    r"|launch__shared_mem_per_block_(dynamic|static)|sm__warps_active\.avg\.pct_of_peak_sustained_active"
    r"|smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed"
    r"|sm__inst_executed_pipe_(alu|fma|fmaheavy|fp64|lsu|xu|uniform)\.avg\.pct_of_peak_sustained_active"
    r"|l1tex__data_pipe_lsu_wavefronts_mem_shared(_op_ld)?\.sum(\.pct_of_peak_sustained_elapsed)?"
    r"|l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_ld)?\.sum|smsp__inst_executed_op_shared_ld\.sum"
    r"|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__thread_inst_executed_per_inst_executed\.ratio"
    r"|smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|sm__cycles_elapsed\.max|smsp__cycles_active\.avg)$")


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"# kernel: {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for h, u, v in zip(hdr, units, r):
            if KEEP.match(h):
                print(f"{h:90s} {v} {u}")


if __name__ == "__main__":
    main()
