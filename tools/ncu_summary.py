#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__cycles_active.avg','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','lts__t_sectors.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','smsp__inst_executed_op_shared_atom.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_op_global_ld.sum','smsp__inst_executed_op_shared_ld.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum']
for r in rows[2:]:
    print("== kernel", r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    for i, h in enumerate(hdr):
        try:
            big = 'issue_stalled' in h and 'per_issue_active' in h and float(r[i] or 0) > 0.3
        except ValueError:
            big = False
        if h in keys or big:
            print(f"  {h} [{units[i]}] = {r[i]}")
