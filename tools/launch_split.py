"""Per-iteration kernel times from an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_split.py launches.csv"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); mi = hdr.index('Metric Name')
seq = [(r[ki].split('(')[0].replace('icpb::', ''), float(r[vi].replace(',', ''))) for r in rows[1:] if r[mi] == 'gpu__time_duration.sum']
names = [n for n, _ in seq]
ends = [i for i, n in enumerate(names) if n.endswith('stage_b_kernel')]
prev = -1
for e in ends:
    it = [(n, t) for n, t in seq[prev + 1:e + 1] if not any(x in n for x in ('radix', 'scan_', 'node_', 'bbox', 'morton', 'gather', 'group_cell', 'group_scatter', 'group_split', 'group_fixed', 'grid_', 'query_keys', 'inv_perm', 'leaf_depth', 'cell_grid', 'root_node', 'cube_root'))]
    print(' | '.join(f"{n[:16]} {t / 1000:.0f}" for n, t in it), ' = %.0f us' % (sum(t for _, t in it) / 1000))
    prev = e
