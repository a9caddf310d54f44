"""Tile-kernel work counters per iteration (profiling aid)."""
import os, sys; sys.path.insert(0, '/root/repo')
if len(sys.argv) > 4: os.environ["ICP_B200_DEBUG_COUNTERS"] = "1"
import numpy as np
from iterativeclosestpoint_b200 import synth
from iterativeclosestpoint_b200.engine import Handle, ICPParameters
m = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
regime = sys.argv[2] if len(sys.argv) > 2 else 'primary'
src, tgt = synth.make_pair(m, 3, regime)
h = Handle(0); h.set_option('nn_mode', int(os.environ.get('ICP_MODE', '3')))
for kv in os.environ.get('ICP_OPTS','').split(','):
    if kv: h.set_option(kv.split('=')[0], float(kv.split('=')[1]))
def cb(it):
    print(f"iter {it.iteration} nn_ms {it.nnMs:.2f} rmse {it.rmse:.4f}", flush=True)
    h.nn_tile_counters()
h.set_callbacks(on_iteration=cb)
h.set_params(ICPParameters(maxIterations=int(sys.argv[3]) if len(sys.argv) > 3 else 14))
w = src.copy(); r = h.register(w, tgt)
info = h.octree_info(); print('nodes', info.n_nodes, 'leaves', info.n_leaves, 'depth', info.depth)
d = h.octree_dump() if m <= 1_000_000 else None
if d is not None:
    lf = d['leaf'].astype(bool); b = d['box'][lf]
    ext = np.stack([b[:,1]-b[:,0], b[:,3]-b[:,2], b[:,5]-b[:,4]],1)
    print('leaf depth hist', np.bincount(d['depth'][lf]))
    print('leaf extent median', np.median(ext,0), 'mean pts/leaf', d['count'][lf].mean())
