#!/usr/bin/env python
"""Per-iteration kernel times from an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_table.py launches.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[h]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
seq = [(r[kn].split('(')[0].replace('icpb::', '').replace('void ', ''), float(r[mv]) / 1e3) for r in rows[h + 1:]
       if len(r) > mv and r[mv].replace('.', '').isdigit()]
line = ''; tot = 0.0
for n, t in seq:
    if not any(k in n for k in ('nn_', 'stat_a', 'stage_b', 'apply_pending', 'solve')):
        continue
    if n.startswith('nn_group') or n.startswith('nn_keep'):
        if line: print(f'{line} | {tot:.0f}us')
        line = ''; tot = 0.0
    line += f'{n[:22]}={t:.0f}  '; tot += t
print(f'{line} | {tot:.0f}us')
