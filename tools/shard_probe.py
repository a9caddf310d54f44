"""One rank's share of a sharded registration on one GPU (profiling aid): 1/parts of the 10M-point source (a spatially compact
shard) against the full target.  usage: shard_probe.py [parts] [m]"""
import os, sys; sys.path.insert(0, '/root/repo')
import numpy as np
from iterativeclosestpoint_b200 import sharding, synth
from iterativeclosestpoint_b200.engine import Handle, ICPParameters
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
src, tgt = synth.make_pair(m, 3, 'primary')
idx = sharding.shard_spatial(src, 0, parts)
shard = np.ascontiguousarray(src[idx])
h = Handle(0)
h.set_params(ICPParameters(maxIterations=12))
r = h.register(shard, tgt)
print(f'shard {len(shard)} of {m}: nn_ms', [round(i.nnMs, 3) for i in r.iterationHistory])
print('   it_ms', [round(i.iterMs, 3) for i in r.iterationHistory])
h.close()
