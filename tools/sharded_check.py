"""Multi-GPU parity check, run under torchrun (one rank per GPU): the sharded registration must equal the single-GPU
registration of the same clouds -- identical iteration count and inlier counts, transforms / RMSE / moved source shards within
north_star's 1e-9 (the deviations found are reported) -- all ranks must hold bit-identical transforms, the resident entry points
must agree with the host ones, a stop request on one rank must cancel every rank, and (ICP_CHECK_ORACLE=1) NN indices of two
seeded samples must equal the oracle's.  Prints one JSON line on rank 0."""
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
src, tgt = sharding.shared_pair(m, 3, "primary", lr, world, dist.barrier)
# ICP_CHECK_SRC=k: only the first k source points against the m-point target (k < world leaves ranks with EMPTY shards)
n_pair = m
m = int(os.environ.get("ICP_CHECK_SRC", m)); src = src[:m]
lo, hi = sharding.shard_range(m, rank, world)
max_it = int(os.environ.get("ICP_CHECK_ITERS", "30"))
with_oracle = os.environ.get("ICP_CHECK_ORACLE", "0") == "1"
h = Handle(lr)
sharding.init_sharded(h, dist, rank, world)
h.set_params(ICPParameters(maxIterations=max_it))
shard = np.ascontiguousarray(src[lo:hi]).copy()
res = h.register_sharded(shard, m, tgt)
# single-GPU reference run on every rank (its own device)
h1 = Handle(lr); h1.set_params(ICPParameters(maxIterations=max_it))
full = src.copy(); ref = h1.register(full, tgt)
ok = True; notes = []
def check(c, msg):
    global ok
    if not c: ok = False; notes.append(msg)
check(res.totalIterations == ref.totalIterations, f"iterations {res.totalIterations} vs {ref.totalIterations}")
check(len(res.iterationHistory) == len(ref.iterationHistory), "history length")
# bars: north_star's 1e-9 on transforms (rank-order versus block-order sums over up to 10^8 terms feed back through the pose:
# 1e-13 .. 4e-12 seen); the deviations actually found go into the JSON line
dev = {"rmse_rel": 0.0, "T_rel": 0.0, "shard_rel": 0.0}
for a, b in zip(res.iterationHistory, ref.iterationHistory):
    check(a.validPoints == b.validPoints, f"valid {a.validPoints} vs {b.validPoints} at {a.iteration}")
    dev["rmse_rel"] = max(dev["rmse_rel"], abs(a.rmse - b.rmse) / b.rmse)
    dev["T_rel"] = max(dev["T_rel"], float(np.max(np.abs(a.transform - b.transform)) / max(1.0, np.max(np.abs(b.transform)))))
if hi > lo: dev["shard_rel"] = float(np.max(np.abs(shard - full[lo:hi])) / np.max(np.abs(full)))
check(dev["rmse_rel"] <= 1e-9, f"rmse rel {dev['rmse_rel']:.3e}")
check(dev["T_rel"] <= 1e-9, f"T rel {dev['T_rel']:.3e}")
check(dev["shard_rel"] <= 1e-9, f"moved shard rel {dev['shard_rel']:.3e}")
# all ranks hold the same bits
t = torch.tensor(res.cumulativeT.reshape(-1), dtype=torch.float64, device="cuda")
g = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(g, t)
check(all(torch.equal(g[0], x) for x in g), "ranks disagree on the cumulative transform bits")
# NN indices against the oracle (the CPU restatement, all host threads) on two seeded samples of rank 0's shard: at the pose
# the run ended in, through this rank's replica of the octree
oracle_note = None
if with_oracle and rank == 0:
    from oracle.binding import Oracle
    orc = Oracle(); otree = orc.octree(tgt)
    rs = np.random.default_rng(44).permutation(hi - lo)
    assert hi - lo >= 60000, "the oracle sample needs 60000 points on rank 0"
    h1.octree_build(tgt, 10, 20)
    bad = 0
    for sample, cloud in ((rs[:30000], np.ascontiguousarray(src[lo:hi])), (rs[30000:60000], shard)):
        q = np.ascontiguousarray(cloud[sample])
        got, _, _ = h1.nn_query(q)
        bad += int(np.count_nonzero(got != otree.find_nearest(q, nthreads=orc.hw_threads())))
    check(bad == 0, f"{bad} NN indices differ from the oracle")
    oracle_note = f"NN indices of 2 x 30000 sampled queries (start pose, final pose) vs the oracle: {bad} mismatches"
# the resident entry points take the same road (redistribution over NVLink, return on write-back)
h.source_upload(np.ascontiguousarray(src[lo:hi]))
back = np.ascontiguousarray(src[lo:hi]).copy()  # (a run that ends without a write-back leaves it as it was, like icp_register_sharded)
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
all_notes = [None] * world; dist.all_gather_object(all_notes, notes[:5])
notes = [f"rank {r}: {n}" for r, ns in enumerate(all_notes) for n in ns]
if rank == 0:
    line = {"sharded_parity_ok": bool(flag.item()), "world": world, "points": m, "iterations": res.totalIterations,
            "final_rmse": res.finalRMSE, "max_deviation_vs_1gpu": dev,
            "checks": ["iterations / inlier counts equal the 1-GPU run", "transforms, rmse, moved shards <= 1e-9 vs 1-GPU (found: max_deviation_vs_1gpu)", "ranks bit-identical", "resident path == host path", "stop on one rank cancels all ranks",
            "run after a cancelled run unharmed"], "timings_ms": {k: round(float(v), 3) for k, v in res.timings_ms.items()},
            "timings_ms_second_run": {k: round(float(v), 3) for k, v in res4.timings_ms.items()}, "oracle": oracle_note, "notes": notes[:16]}
    print(json.dumps(line), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    line["target_points"] = n_pair
    with open(f"gpurun_out/sharded_check_{world}gpu_{m}.json", "w") as f:
        json.dump(line, f)
dist.barrier(); dist.destroy_process_group(); h.close(); h1.close()
sys.exit(0 if flag.item() else 1)
