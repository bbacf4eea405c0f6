#!/usr/bin/env python
"""Top warp-stall reasons (PC sampling) and pipe utilisation of each kernel in an .ncu-rep."""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[0]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:90])
    def g(k):
        return r[hdr.index(k)] if k in hdr else "n/a"
    for k in ("gpu__time_duration.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
              "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
              "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_cbu.sum",
              "sm__inst_executed_pipe_adu.sum", "sm__inst_executed_pipe_uniform.sum", "smsp__inst_issued.sum", "sm__cycles_active.avg",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "launch__registers_per_thread"):
        print(f"   {k:70s} {g(k)}")
    st = [(float(r[i].replace(',', '') or 0), h) for i, h in enumerate(hdr)
          if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and r[i] not in ('', 'n/a')]
    tot = sum(v for v, _ in st) or 1
    for v, h in sorted(st, reverse=True)[:9]:
        print('      %6.1f%% %s' % (100 * v / tot, h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
