"""CPU tests of the N > 1 host logic with torch.distributed (gloo, world_size 2, 127.0.0.1): shard ranges, the exchange
that ships rank 0's id to everyone, and the rank-ordered merge of the per-rank partial records -- the same reduction
the library performs on the device after its two NCCL all-gathers.  Every rank must end up with
bit-identical statistics, equal to the single-process result within the stated tolerance, and the inlier mask and
transform derived from them must equal the oracle's for the whole cloud."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from iterativeclosestpoint_b200 import sharding, synth  # noqa: E402
import shard_merge  # noqa: E402  (tests/shard_merge.py)


def test_shard_ranges_tile_the_source():
    for n in (0, 1, 7, 1000, 10_000_019):
        for world in (1, 2, 3, 8):
            cuts = [sharding.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_block_cyclic_ranges_tile_the_source_and_round_trip():
    r = np.random.default_rng(4)
    for n in (0, 1, 7, 1000, 300_017):
        for world in (1, 2, 3, 8):
            for block in (1, 64, 65536):
                per_rank = [sharding.shard_blocks(n, k, world, block) for k in range(world)]
                flat = sorted(x for rr in per_rank for x in rr)
                assert (flat[0][0] == 0 and flat[-1][1] == n) if n else not flat
                assert all(a[1] == b[0] for a, b in zip(flat, flat[1:]))
                sizes = [sum(hi - lo for lo, hi in rr) for rr in per_rank]
                assert max(sizes) - min(sizes) <= block
    pts = r.normal(size=(1000, 3))
    out = np.zeros_like(pts)
    for k in range(3):
        rr = sharding.shard_blocks(len(pts), k, 3, 64)
        sh = sharding.take_shard(pts, rr)
        assert sh.flags["C_CONTIGUOUS"] and len(sh) == sum(hi - lo for lo, hi in rr)
        sharding.put_shard(out, rr, sh)
    assert np.array_equal(out, pts)
    with pytest.raises(ValueError):
        sharding.shard_blocks(10, 0, 2, 0)


def test_spatial_shards_partition_the_cloud_into_compact_equal_parts():
    pts = synth.make_target(40_000, 5)
    order = sharding.spatial_order(pts)
    assert sorted(order.tolist()) == list(range(len(pts)))
    world = 8
    parts = [sharding.shard_spatial(pts, k, world, order) for k in range(world)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(len(pts)))
    assert max(map(len, parts)) - min(map(len, parts)) <= 1
    # compact: a shard's bounding box covers far less than the cloud's (a range of the random caller order covers all of it)
    area = lambda p: float(np.prod((p.max(0) - p.min(0))[:2]))
    assert max(area(pts[i]) for i in parts) < 0.5 * area(pts) and np.mean([area(pts[i]) for i in parts]) < 0.3 * area(pts)
    assert len(sharding.spatial_order(np.empty((0, 3)))) == 0


def test_rank_ordered_merge_matches_whole_cloud_statistics():
    r = np.random.default_rng(0)
    d = np.abs(r.normal(size=100_003)) * 0.3
    whole = shard_merge.stat_partial(d)
    for world in (2, 3, 8):
        parts = [shard_merge.stat_partial(d[slice(*sharding.shard_range(len(d), k, world))]) for k in range(world)]
        m = shard_merge.merge_in_rank_order(parts)
        assert m[0] == whole[0] and m[3] == whole[3] and m[4] == whole[4]
        assert abs(m[1] - whole[1]) <= 1e-14 * whole[1] and abs(m[2] - whole[2]) <= 1e-12 * whole[2]


def test_rank_ordered_merge_with_more_ranks_than_points():
    """A world larger than the cloud leaves ranks with EMPTY shards (their partial has n = 0); the merge in rank order must
    still give the whole cloud's statistics (the device path: tools/sharded_check.py with ICP_CHECK_SRC, profiles/r2_sharded_check_8gpu_21pts.json)."""
    d = np.array([0.3, 0.1, 0.25, 0.7, 0.05])
    whole = shard_merge.stat_partial(d)
    for world in (8, 16):
        cuts = [sharding.shard_range(len(d), k, world) for k in range(world)]
        assert sum(1 for a, b in cuts if a == b) == world - len(d)
        parts = [shard_merge.stat_partial(d[a:b]) for a, b in cuts]
        m = shard_merge.merge_in_rank_order(parts)
        assert m[0] == whole[0] and m[3] == whole[3] and m[4] == whole[4]
        assert abs(m[1] - whole[1]) <= 1e-15 and abs(m[2] - whole[2]) <= 1e-15


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from iterativeclosestpoint_b200 import sharding as sh, synth as sy
    from oracle.binding import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        src, tgt = sy.make_pair(6000, 2, "primary")
        orc = Oracle()
        lo, hi = sh.shard_range(len(src), rank, world)
        # the id exchange: a stand-in object with the handle's interface (no GPU here)
        class FakeHandle:
            def comm_unique_id(self):
                return bytes(range(128))
        uid = sh.exchange_unique_id(FakeHandle(), dist, rank)
        assert uid == bytes(range(128))
        # this rank's shard: NN (oracle stands in for the device kernel), distances, stage-A partial
        idx = orc.octree(tgt).find_nearest(src[lo:hi])
        dv = src[lo:hi] - tgt[idx]
        d = np.sqrt(dv[:, 0] * dv[:, 0] + dv[:, 1] * dv[:, 1] + dv[:, 2] * dv[:, 2])
        gathered = [None] * world
        dist.all_gather_object(gathered, shard_merge.stat_partial(d))
        merged = shard_merge.merge_in_rank_order(gathered)
        mean, std, thr = shard_merge.threshold(merged, len(src), 3.0, 0)
        pa = pb = (tgt.min(0) + tgt.max(0)) * 0.5
        gathered_b = [None] * world
        dist.all_gather_object(gathered_b, shard_merge.moment_partial(src[lo:hi], tgt[idx], d, thr, pa, pb))
        mom = shard_merge.sum_in_rank_order(gathered_b)
        cA, cB, H = shard_merge.moments_to_H(mom, pa, pb)
        T = orc.solve_from_H(H, cA, cB)
        q.put((rank, merged.tobytes(), float(thr), mom.tobytes(), T.tobytes(), (d <= thr).astype(np.uint8).tobytes(), lo, hi))
    finally:
        dist.destroy_process_group()


def test_two_ranks_agree_bitwise_and_match_the_oracle():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank derived bit-identical statistics, moments and transform
    assert out[0][1] == out[1][1] and out[0][2] == out[1][2] and out[0][3] == out[1][3] and out[0][4] == out[1][4]
    # and they are the whole-cloud answer of the oracle
    from oracle.binding import Oracle
    src, tgt = synth.make_pair(6000, 2, "primary")
    orc = Oracle()
    idx = orc.octree(tgt).find_nearest(src)
    odist, omask, ost = orc.iteration_stats(src, tgt, idx, 0, 3.0, 0)
    thr = out[0][2]
    assert abs(thr - ost.threshold) <= 1e-12 * ost.threshold
    mask = np.concatenate([np.frombuffer(o[5], dtype=np.uint8) for o in out])
    assert np.array_equal(mask, omask)
    T = np.frombuffer(out[0][4]).reshape(4, 4)
    oT = orc.kabsch(src[omask.astype(bool)], tgt[idx[omask.astype(bool)]])
    assert np.max(np.abs(T - oT)) <= 1e-9
