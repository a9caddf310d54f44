"""BASELINE.json config #4 at full size on one GPU (100 M <-> 100 M, ~2.4 GB per cloud): NN indices against the oracle on two
disjoint 30k samples at the starting pose, the default search mode against the one-thread-per-query walk over whole runs
(bit-identical transforms), and a sample check after the run.  Prints one JSON line.  usage: config4_check.py [m]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iterativeclosestpoint_b200 import synth
from iterativeclosestpoint_b200.engine import Handle, ICPParameters
from oracle.binding import Oracle

m = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
t0 = time.time()
src, tgt = synth.make_pair(m, 4, "primary")
t_gen = time.time() - t0
orc = Oracle()
t0 = time.time()
otree = orc.octree(tgt)
t_tree = time.time() - t0
h = Handle(0)
h.octree_build(tgt, 10, 20)
idx, dist, nn_ms = h.nn_query(src)
r = np.random.default_rng(44).permutation(m)
ok = True
notes = []
for k, sample in enumerate((r[:30000], r[30000:60000])):
    want = otree.find_nearest(src[sample], nthreads=orc.hw_threads())
    if not np.array_equal(idx[sample], want):
        ok = False
        notes.append(f"sample {k}: {int((idx[sample] != want).sum())} indices differ")
dv = src[r[:200000]] - tgt[idx[r[:200000]]]
if not np.array_equal(dist[r[:200000]], np.sqrt(dv[:, 0] * dv[:, 0] + dv[:, 1] * dv[:, 1] + dv[:, 2] * dv[:, 2])):
    ok = False
    notes.append("distances are not sqrt of the matched pair's sum of squares")
runs = {}
for mode in (3, 6):
    h.set_option("nn_mode", mode)
    h.set_params(ICPParameters(maxIterations=8, tolerance=1e-15))
    work = src.copy()
    runs[mode] = h.register(work, tgt)
    if mode == 6:
        moved = work
    else:
        del work
a, b = runs[3], runs[6]
if not (a.loopIterations == b.loopIterations and np.array_equal(a.cumulativeT, b.cumulativeT)
        and [x.rmse for x in a.iterationHistory] == [x.rmse for x in b.iterationHistory]
        and [x.validPoints for x in a.iterationHistory] == [x.validPoints for x in b.iterationHistory]):
    ok = False
    notes.append("default mode and per-thread walk disagree")
h.octree_build(tgt, 10, 20)
idx2, _, _ = h.nn_query(moved)
sample = r[60000:90000]
if not np.array_equal(idx2[sample], otree.find_nearest(moved[sample], nthreads=orc.hw_threads())):
    ok = False
    notes.append("indices after the run differ from the oracle")
print(json.dumps({"config4_parity_ok": ok, "points": m, "iterations": b.loopIterations,
                  "rmse": [x.rmse for x in b.iterationHistory], "nn_ms_default": [round(x.nnMs, 2) for x in b.iterationHistory],
                  "nn_ms_per_thread_walk": [round(x.nnMs, 2) for x in a.iterationHistory], "stateless_nn_query_ms": nn_ms,
                  "host_s": {"generate": round(t_gen, 1), "oracle_octree": round(t_tree, 1)}, "notes": notes}), flush=True)
h.close()
sys.exit(0 if ok else 1)
