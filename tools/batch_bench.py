"""BASELINE.json config #5 (many small independent registrations): pairs/s through icp_register_batch vs the
compiled reference on all host threads (profiling aid; prints one JSON line)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from iterativeclosestpoint_b200 import synth
from iterativeclosestpoint_b200.engine import Handle, ICPParameters
n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 512
pts = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
workers = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,8,16").split(",")]
pairs = [synth.small_pair(p, n=pts) for p in range(n_pairs)]
out = {"pairs": n_pairs, "points": pts}
h = Handle(0); h.set_params(ICPParameters())
for w in workers:
    h.set_option("batch_workers", w)
    srcs = [s.copy() for s, _ in pairs]
    h.register_batch(srcs[:16], [t for _, t in pairs[:16]])  # warm-up (worker creation, allocations)
    srcs = [s.copy() for s, _ in pairs]
    t0 = time.perf_counter(); res = h.register_batch(srcs, [t for _, t in pairs]); dt = time.perf_counter() - t0
    out[f"gpu_pairs_per_s_w{w}"] = n_pairs / h.last_batch_seconds
    out[f"gpu_pairs_per_s_w{w}_incl_python"] = n_pairs / dt
    out["mean_iterations"] = float(np.mean([r.totalIterations for r in res]))
try:
    from oracle import binding
    if binding.ref_available():
        ref = binding.RefEngine(); nthr = os.cpu_count() or 1
        sub = pairs[: min(n_pairs, 4 * nthr)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(nthr) as ex:
            list(ex.map(lambda st: ref.icp(st[0], st[1]).total_iterations, sub))
        out["cpu_reference_pairs_per_s"] = len(sub) / (time.perf_counter() - t0); out["cpu_threads"] = nthr
except Exception as e:  # noqa
    out["cpu_error"] = str(e)
print(json.dumps(out))
