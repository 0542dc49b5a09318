#!/usr/bin/env python3
"""Summarise an `ncu --page raw --csv` export: one block of key metrics per profiled launch.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_extract.py raw.csv"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second']

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('=====', r[idx['Kernel Name']][:110])
    for w in WANT:
        if w in idx:
            print('  %-72s %16s %s' % (w, r[idx[w]], units[idx[w]]))
    st = []
    for h in hdr:
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
            try:
                st.append((float(r[idx[h]].replace(',', '')), h))
            except ValueError:
                pass
    print('  top stall reasons (warps stalled per issue-active cycle):')
    for v, h in sorted(st, reverse=True)[:7]:
        print('     %7.2f  %s' % (v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
