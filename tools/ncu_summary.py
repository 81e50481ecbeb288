#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers the roofline discussion needs.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print("kernel:", r[h.index("Kernel Name")][:160])
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print(f"  {k:85s} {r[i]:>18s} {units[i]}")
        stalls = {}
        for i, name in enumerate(h):
            m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", name)
            if m:
                try:
                    stalls[m.group(1)] = float(r[i])
                except ValueError:
                    pass
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        print("  top stall reasons (warps stalled per issue-active cycle):", ", ".join(f"{k}={v:.2f}" for k, v in top))


if __name__ == "__main__":
    main(sys.argv[1])
