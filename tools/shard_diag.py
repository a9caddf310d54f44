"""Diagnostic for tiny shards under torchrun: per-iteration records of the sharded run next to the 1-GPU run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from iterativeclosestpoint_b200 import sharding, synth
from iterativeclosestpoint_b200.engine import Handle, ICPParameters
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
h = Handle(lr); sharding.init_sharded(h, dist, rank, world)
h1 = Handle(lr)
for m in [int(x) for x in sys.argv[1].split(",")]:
    for redis in (1, 0):
        src, tgt = synth.make_pair(max(m, 3), 3, "primary")
        src = src[:m]
        lo, hi = sharding.shard_range(m, rank, world)
        h.set_option("redistribute", redis)
        h.set_params(ICPParameters(maxIterations=12)); h1.set_params(ICPParameters(maxIterations=12))
        shard = np.ascontiguousarray(src[lo:hi]).copy()
        res = h.register_sharded(shard, m, tgt)
        full = src.copy(); ref = h1.register(full, tgt)
        if rank == 0:
            print(f"m={m} redistribute={redis}: sharded it={res.totalIterations} status={res.status} ref it={ref.totalIterations} status={ref.status}")
            for a, b in zip(res.iterationHistory, ref.iterationHistory):
                print(f"   it {a.iteration}: valid {a.validPoints}/{b.validPoints} rmse {a.rmse:.12g} / {b.rmse:.12g}")
        dist.barrier()
dist.destroy_process_group()
