"""Multi-GPU parity check, run under torchrun (one rank per GPU): the sharded registration must equal the single-GPU
registration of the same clouds -- identical iteration count and inlier counts, transforms to 1e-12, RMSE to 1e-10 (rank-order
versus block-order sums feed back through the pose: 1.2e-12 seen at 4 ranks x 4M points where the RMSE falls fastest; the
end-to-end bar is 1e-9), moved source shards equal to 1e-12 -- and all ranks must hold bit-identical transforms.  Prints one JSON line on rank 0."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from iterativeclosestpoint_b200 import sharding, synth
from iterativeclosestpoint_b200.engine import Handle, ICPParameters

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
m = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
src, tgt = synth.make_pair(m, 3, "primary")
lo, hi = sharding.shard_range(m, rank, world)
h = Handle(lr)
sharding.init_sharded(h, dist, rank, world)
h.set_params(ICPParameters(maxIterations=30))
shard = np.ascontiguousarray(src[lo:hi]).copy()
res = h.register_sharded(shard, m, tgt)
# single-GPU reference run on every rank (its own device)
h1 = Handle(lr); h1.set_params(ICPParameters(maxIterations=30))
full = src.copy(); ref = h1.register(full, tgt)
ok = True; notes = []
def check(c, msg):
    global ok
    if not c: ok = False; notes.append(msg)
check(res.totalIterations == ref.totalIterations, f"iterations {res.totalIterations} vs {ref.totalIterations}")
check(len(res.iterationHistory) == len(ref.iterationHistory), "history length")
for a, b in zip(res.iterationHistory, ref.iterationHistory):
    check(a.validPoints == b.validPoints, f"valid {a.validPoints} vs {b.validPoints} at {a.iteration}")
    check(abs(a.rmse - b.rmse) <= 1e-10 * b.rmse, f"rmse at {a.iteration}: rel {abs(a.rmse - b.rmse) / b.rmse:.3e} outliers {a.outlierPoints} vs {b.outlierPoints}")
    check(np.max(np.abs(a.transform - b.transform)) <= 1e-12 * max(1.0, np.max(np.abs(b.transform))), f"T at {a.iteration}")
check(np.max(np.abs(shard - full[lo:hi])) <= 1e-12 * np.max(np.abs(full)), "moved shard")
# all ranks hold the same bits
t = torch.tensor(res.cumulativeT.reshape(-1), dtype=torch.float64, device="cuda")
g = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(g, t)
check(all(torch.equal(g[0], x) for x in g), "ranks disagree on the cumulative transform bits")
# the resident entry points take the same road (redistribution over NVLink, return on write-back)
h.source_upload(np.ascontiguousarray(src[lo:hi]))
back = np.zeros((hi - lo, 3))
res2 = h.register_resident(m, source_out=back)
check(res2.totalIterations == ref.totalIterations and np.array_equal(res2.cumulativeT, res.cumulativeT), "resident run differs from icp_register_sharded")
check(np.array_equal(back, shard), "resident write-back differs")
# a stop request on ONE rank must end the run on EVERY rank in the same iteration, sources untouched (icpengine.cpp:160-164)
import ctypes
stop = ctypes.c_int(1 if rank == world - 1 else 0)
shard3 = np.ascontiguousarray(src[lo:hi]).copy()
res3 = h.register_sharded(shard3, m, tgt, stop_flag=stop)
check(res3.status == 2 and not res3.success and res3.totalIterations == 0, f"stop: status {res3.status} iterations {res3.totalIterations}")
check(np.array_equal(shard3, src[lo:hi]), "stop: the source was touched")
# ... and the next run on the same handles is unharmed (epochs keyed on the run)
shard4 = np.ascontiguousarray(src[lo:hi]).copy()
res4 = h.register_sharded(shard4, m, tgt)
check(np.array_equal(res4.cumulativeT, res.cumulativeT) and np.array_equal(shard4, shard), "run after a cancelled run differs")
flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    line = {"sharded_parity_ok": bool(flag.item()), "world": world, "points": m, "iterations": res.totalIterations,
            "final_rmse": res.finalRMSE, "checks": ["iterations / inlier counts equal the 1-GPU run", "transforms <= 1e-12, rmse <= 1e-10 vs 1-GPU",
            "moved shards <= 1e-12 vs 1-GPU", "ranks bit-identical", "resident path == host path", "stop on one rank cancels all ranks",
            "run after a cancelled run unharmed"], "timings_ms": {k: round(float(v), 3) for k, v in res.timings_ms.items()}, "notes": notes[:5]}
    print(json.dumps(line), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/sharded_check_{world}gpu_{m}.json", "w") as f:
        json.dump(line, f)
dist.barrier(); dist.destroy_process_group(); h.close(); h1.close()
sys.exit(0 if flag.item() else 1)
