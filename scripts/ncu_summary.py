#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / profiles/ quote.
usage: scripts/ncu_summary.py <report.ncu-rep> [--source]"""
import csv, io, subprocess, sys, collections

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u = rows[0], rows[1]
    for d in rows[2:]:
        print("== kernel", d[h.index("Kernel Name")] if "Kernel Name" in h else "?")
        for i, name in enumerate(h):
            if name in KEYS or name.startswith("smsp__average_warps_issue_stalled") and name.endswith("per_issue_active.ratio"):
                print("%-86s %-10s %s" % (name, u[i], d[i]))


def source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] in ("#", "Line #", "Address")]
    if not hi:
        print("no source view"); return
    h = rows[hi[0]]
    data = rows[hi[0] + 1:]
    si = h.index("Source"); sa = h.index("Warp Stall Sampling (All Samples)"); ie = h.index("Instructions Executed")
    tot_s = sum(int(r[sa] or 0) for r in data if len(r) > sa); tot_i = sum(int(r[ie] or 0) for r in data if len(r) > ie)
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    top = sorted((r for r in data if len(r) > sa), key=lambda r: -int(r[ie] or 0))[:45]
    for r in top:
        print("%5.1f%% inst %5.1f%% samp | %s" % (100 * int(r[ie] or 0) / max(1, tot_i), 100 * int(r[sa] or 0) / max(1, tot_s), r[si].strip()[:150]))


if __name__ == "__main__":
    raw(sys.argv[1])
    if "--source" in sys.argv:
        source(sys.argv[1])
