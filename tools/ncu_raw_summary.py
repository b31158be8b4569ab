#!/usr/bin/env python
"""Key metrics of an `ncu -i X.ncu-rep --page raw --csv` dump (last kernel row). usage: ncu_raw_summary.py raw.csv [title]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h, u, v = rows[0], rows[1], rows[-1]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_inst_executed_op_tma_ld.sum"]
keep += [x for x in h if x.startswith("smsp__average_warps_issue_stalled") and x.endswith("per_issue_active.ratio")]
print(sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none")
name = [i for i, x in enumerate(h) if x == "Kernel Name"]
if name:
    print("kernel =", v[name[0]])
for k in keep:
    if k in h:
        print("%s = %s %s" % (k, v[h.index(k)], u[h.index(k)]))
