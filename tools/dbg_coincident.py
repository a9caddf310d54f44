import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, clouds
from iterativeclosestpoint_b200.engine import Handle, ICPParameters
tgt = clouds.coincident(200)
for mode in (0,1,3,2):
    h = Handle(0)
    h.set_option("nn_mode", mode)
    h.octree_build(tgt, 10, 20)
    t = h.lib  # noqa
    info = h.octree_info()
    print("mode", mode, "built", info.n_nodes, flush=True)
    for qname, q in clouds.query_sets(tgt).items():
        try:
            idx, _, _ = h.nn_query(q[:1500])
            print("  ", qname, "ok", idx[:3], flush=True)
        except Exception as e:
            print("  ", qname, "FAILED", e, flush=True); break
    h.close()
