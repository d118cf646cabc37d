#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the text summary kept under profiles/: per kernel, the metrics the
DESIGN.md analysis quotes.  usage: python tools/ncu_summary.py report.ncu-rep "header line" > profiles/xyz.txt"""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")


def main():
    rep, header = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    names, units = rows[0], rows[1]
    print(header)
    for r in rows[2:]:
        d = dict(zip(names, r))
        u = dict(zip(names, units))
        print("== %s" % d["Kernel Name"])
        for k in KEEP:
            if k in d:
                print("   %s [%s] %s" % (k, u[k], d[k]))
        for k in names:
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
                print("   %s [%s] %s" % (k, u[k], d[k]))


if __name__ == "__main__":
    main()
