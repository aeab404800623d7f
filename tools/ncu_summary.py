#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) into the handful of numbers the roofline discussion needs.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/rNN_<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if sub and sub not in r[ki]:
            continue
        print(f"== {r[ki]}  (launch id {r[0]})")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:75s} {r[i]:>18s} {units[i]}")
        stalls = [(h, r[i]) for i, h in enumerate(hdr) if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")]
        stalls = sorted(stalls, key=lambda kv: -float(kv[1].replace(",", "") or 0))[:8]
        print("  top stall reasons (warps stalled per issue-active cycle):")
        for h, v in stalls:
            print(f"    {h[len(STALL):-len('_per_issue_active.ratio')]:40s} {v}")
        print()


if __name__ == "__main__":
    main()
