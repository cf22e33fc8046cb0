#!/usr/bin/env python3
"""Summarise an `ncu --set full` report (.ncu-rep) into the text committed under profiles/.
usage: summarize_ncu.py report.ncu-rep [--source]  (needs the `ncu` CLI, no GPU)"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum']
rep = sys.argv[1]
rows = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
H = rows[0]
for r in rows[2:]:
    print('---')
    print('  %-84s %s' % ('Kernel Name', r[H.index('Kernel Name')]))
    for k in KEYS:
        if k in H:
            print('  %-84s %s %s' % (k, r[H.index(k)], rows[1][H.index(k)]))
    st = [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(r[i])) for i, h in enumerate(H)
          if 'issue_stalled' in h and 'per_issue_active' in h]
    print('  stall cycles per issued instruction: ' + ', '.join('%s %.2f' % x for x in sorted(st, key=lambda x: -x[1])[:8]))
