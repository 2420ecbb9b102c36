#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: per kernel the duration, DRAM bytes, pipe
utilisation and the top warp-stall reasons.  usage: ncu_summary.py prof_raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
H, U = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum']
for r in rows[2:]:
    print('==== %s  (id %s)' % (r[H.index('Kernel Name')][:50], r[0]))
    for w in want:
        if w in H:
            i = H.index(w)
            print('  %-72s %s %s' % (w, r[i], U[i]))
    vals = []
    for i, h in enumerate(H):
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
            try:
                vals.append((float(r[i].replace(',', '')), h))
            except ValueError:
                pass
    for v, h in sorted(vals, reverse=True)[:7]:
        print('   stall %-28s %.3f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
