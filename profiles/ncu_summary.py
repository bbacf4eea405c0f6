#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per-kernel duration, DRAM traffic, issue/occupancy
numbers and the top warp-stall reasons.  Usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    stalls = [(h, i) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_warp_active.pct")]
    for r in rows[2:]:
        print("=" * 100)
        print(r[name_i][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:66s} {r[i]:>18s} {units[i]}")
        top = sorted(((float(r[i].replace(',', '') or 0), h) for h, i in stalls), reverse=True)[:7]
        for v, h in top:
            print(f"  stall {v:7.2f}%  {h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', '')}")


if __name__ == "__main__":
    main(sys.argv[1])
