"""Per-iteration NN times of nn_mode 5 (keep + search) over K / alpha / bias (profiling aid).  usage: keep_sweep.py [m] [regime]"""
import os, sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
from iterativeclosestpoint_b200 import synth
from iterativeclosestpoint_b200.engine import Handle, ICPParameters
m = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
regime = sys.argv[2] if len(sys.argv) > 2 else 'primary'
src, tgt = synth.make_pair(m, 3, regime)
combos = [(4, None, None, None, None)] + [(5, k, a, b, rc) for (k, a, b, rc) in
          [(4, 2.0, 1, 0.45), (4, 2.0, 0, 0.45), (4, 1.5, 0, 0.45), (4, 3.0, 0, 0.45), (4, 2.0, 0, 0.3), (4, 2.0, 0, 0.6), (2, 2.0, 0, 0.45), (3, 2.0, 0, 0.45)]]
if os.environ.get("KEEP_COMBOS"):
    combos = [tuple(float(x) if '.' in x else int(x) for x in c.split(':')) for c in os.environ["KEEP_COMBOS"].split(',')]
for mode, k, a, b, rc in combos:
    h = Handle(0); h.set_option('nn_mode', mode); h.set_option('count', float(os.environ.get('ICP_COUNT', '0')))
    if mode == 5:
        h.set_option('keep_k', k); h.set_option('keep_alpha', a); h.set_option('keep_bias', b); h.set_option('keep_rcap', rc)
    h.set_params(ICPParameters(maxIterations=int(os.environ.get("ICP_ITERS", "16"))))
    w = src.copy(); r = h.register(w, tgt)
    nn = [i.nnMs for i in r.iterationHistory]
    print(f'{regime} m={m} mode={mode} k={k} alpha={a} bias={b} rcap={rc} iters={r.loopIterations} nn_total={sum(nn):.2f} nn[3:13]={sum(nn[3:13]):.2f} loop={r.timings_ms["loop"]:.2f}')
    print('   nn_ms:', [round(x, 2) for x in nn])
    if os.environ.get('ICP_COUNT'):
        print('   counters', h.nn_counters(), 'tile', h.nn_tile_counters())
    h.close()
