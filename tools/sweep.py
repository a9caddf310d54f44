"""Option sweep: total / per-iteration NN time of one registration per option set (profiling aid).
usage: sweep.py <points> <regime> <iters> "opt=val,opt=val" ["..." ...]"""
import os, sys; sys.path.insert(0, '/root/repo')
import numpy as np
from iterativeclosestpoint_b200 import synth
from iterativeclosestpoint_b200.engine import Handle, ICPParameters
m = int(sys.argv[1]); regime = sys.argv[2]; iters = int(sys.argv[3])
src, tgt = synth.make_pair(m, 3, regime)
for opts in sys.argv[4:]:
    h = Handle(0)
    for kv in opts.split(','):
        if kv and kv != '-': h.set_option(kv.split('=')[0], float(kv.split('=')[1]))
    h.set_params(ICPParameters(maxIterations=iters))
    w = src.copy(); r = h.register(w, tgt)
    nn = [i.nnMs for i in r.iterationHistory]; it = [i.iterMs for i in r.iterationHistory]
    print(f"{opts:60s} build {r.timings_ms['build']:6.1f}  nn_sum {sum(nn):7.2f}  iter_sum {sum(it):7.2f}  nn[0,2,6,-1] = {nn[0]:.2f} {nn[2]:.2f} {nn[6]:.2f} {nn[-1]:.2f}  n={len(nn)}", flush=True)
    i = h.octree_info(); print(f"      search nodes {i.search_nodes} depth {i.search_depth} grid levels {i.grid_base_level}..{i.grid_fine_level} base cell {i.grid_base_cell:.3f} m grid {i.grid_bytes/1e6:.0f} MB", flush=True)
    h.close()
