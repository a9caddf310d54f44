"""Bring-up probe of the box search (mode 7): one registration with the given options, progress printed per iteration.
usage: box_probe.py <points> <regime> [key=value ...]"""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from iterativeclosestpoint_b200 import synth
from iterativeclosestpoint_b200.engine import Handle, ICPParameters
m = int(sys.argv[1]); regime = sys.argv[2]
src, tgt = synth.make_pair(m, 3, regime)
h = Handle(0); h.set_option('nn_mode', 7)
for kv in sys.argv[3:]:
    k, v = kv.split('='); h.set_option(k, float(v))
h.set_params(ICPParameters(maxIterations=12))
w = src.copy()
print('start', sys.argv[1:], flush=True)
r = h.register(w, tgt)
print('iters', r.loopIterations, 'nn_ms', [round(i.nnMs, 3) for i in r.iterationHistory], flush=True)
print('rmse', [round(i.rmse, 4) for i in r.iterationHistory], flush=True)
h.set_option('nn_mode', 3)
w2 = src.copy(); r2 = h.register(w2, tgt)
print('same as walk:', np.array_equal(r.cumulativeT, r2.cumulativeT), [round(i.nnMs, 3) for i in r2.iterationHistory], flush=True)
h.close()
