#!/usr/bin/env python
"""Prints the key metrics and stall breakdown of every kernel in an .ncu-rep (read with `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "smsp__inst_executed_op_shared_ld.sum"]
for r in rows[2:]:
    print("----")
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:75s} {r[i]} {units[i]}")
    st = [(float(r[i].replace(",", "")), k) for i, k in enumerate(hdr)
          if "smsp__average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio") and r[i]]
    for v, k in sorted(st, reverse=True)[:9]:
        print(f"   stall {v:7.3f} {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")
