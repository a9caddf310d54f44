"""BASELINE.json config #4 size on one GPU, quick form: the first iterations of a registration must reproduce the RMSE history
recorded by tools/config4_check.py (profiles/r05_config4_100m_parity.json, checked against the oracle there) within north_star's
1e-9 (the sums of stages A and B are grouped differently since that record was taken, so the last digits may differ).
usage: config4_quick.py [iterations]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterativeclosestpoint_b200 import synth
from iterativeclosestpoint_b200.engine import Handle, ICPParameters
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
want = json.load(open(os.path.join(root, "profiles", "r05_config4_100m_parity.json")))
k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
src, tgt = synth.make_pair(want["points"], 4, "primary")
h = Handle(0)
h.set_params(ICPParameters(maxIterations=k, tolerance=1e-15))
res = h.register(src, tgt)
got = [x.rmse for x in res.iterationHistory]
info = h.octree_info()
dev = max(abs(a - b) / b for a, b in zip(got, want["rmse"][:k]))
print(json.dumps({"config4_quick_ok": len(got) == k and dev <= 1e-9, "max_rel_deviation": dev, "points": want["points"], "rmse": got, "recorded": want["rmse"][:k],
                  "octree": {"nodes": int(info.n_nodes), "leaves": int(info.n_leaves), "depth": int(info.depth)},
                  "timings_ms": {a: round(float(b), 2) for a, b in res.timings_ms.items()}}), flush=True)
h.close()
