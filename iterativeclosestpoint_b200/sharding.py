"""Host-side plumbing of the multi-GPU path: the source cloud is sharded by contiguous point range,
the target octree is replicated, one process drives one GPU.  `torch.distributed` is used for exactly two things: to
hand rank 0's NCCL unique id to the other ranks, and (in tests on CPU, gloo) to stand in for the library's own two
per-iteration all-gathers.  The data path itself never goes through torch.

(The numpy restatements of the device's record merges that the gloo tests use live with those tests: tests/shard_merge.py.)
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous slice of n source points; slices tile [0, n) and differ by at most one point."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank/world out of range")
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_blocks(n: int, rank: int, world: int, block: int = 65536) -> list[tuple[int, int]]:
    """Block-cyclic point ranges: [0, n) cut into contiguous blocks of `block` points, dealt round-robin to the ranks.
    Still sharding by point range, but a cloud whose caller order is spatially sorted (scan strips, or the synthetic
    scene's terrain-then-boxes order) no longer hands one rank all the hard queries: the per-iteration time is the
    maximum over ranks.  The ranges of all ranks tile [0, n)."""
    if world <= 0 or not (0 <= rank < world) or block <= 0:
        raise ValueError("rank/world/block out of range")
    return [(lo, min(lo + block, n)) for lo in range(rank * block, n, world * block)]


def take_shard(points: np.ndarray, ranges) -> np.ndarray:
    """The rank's shard as one C-contiguous (k, 3) array: its ranges concatenated in order."""
    if not ranges:
        return np.empty((0, 3), dtype=np.float64)
    return np.ascontiguousarray(np.concatenate([points[lo:hi] for lo, hi in ranges]))


def put_shard(points: np.ndarray, ranges, shard: np.ndarray) -> None:
    """Write a (moved) shard back into the caller's array, range by range (inverse of take_shard)."""
    at = 0
    for lo, hi in ranges:
        points[lo:hi] = shard[at:at + (hi - lo)]
        at += hi - lo


def _spread10(v: np.ndarray) -> np.ndarray:
    """10-bit integers -> bits at positions 0, 3, 6, ... (one axis of a 30-bit Morton key)."""
    v = v.astype(np.uint32) & np.uint32(0x3FF)
    v = (v | (v << np.uint32(16))) & np.uint32(0x030000FF)
    v = (v | (v << np.uint32(8))) & np.uint32(0x0300F00F)
    v = (v | (v << np.uint32(4))) & np.uint32(0x030C30C3)
    v = (v | (v << np.uint32(2))) & np.uint32(0x09249249)
    return v


def spatial_order(points: np.ndarray) -> np.ndarray:
    """Stable argsort of the points by a 30-bit Morton key over their own bounding cube (host side, once per cloud)."""
    pts = np.asarray(points, dtype=np.float64)
    if len(pts) == 0:
        return np.empty(0, dtype=np.int64)
    lo = pts.min(axis=0)
    cube = float((pts.max(axis=0) - lo).max())
    scale = 1023.999 / cube if cube > 0.0 else 0.0
    q = np.clip((pts - lo) * scale, 0, 1023).astype(np.uint32)
    key = _spread10(q[:, 0]) | (_spread10(q[:, 1]) << np.uint32(1)) | (_spread10(q[:, 2]) << np.uint32(2))
    return np.argsort(key, kind="stable")


def shard_spatial(points: np.ndarray, rank: int, world: int, order: np.ndarray | None = None) -> np.ndarray:
    """Indices (ascending) of rank's shard when the cloud is first put in spatial (Morton) order and THEN cut into `world`
    contiguous point ranges: every rank owns a compact region with the same number of points, so its queries touch only
    that region's part of the replicated target structure (on 8 GPUs the NN stage of a randomly ordered 10 M-point cloud
    runs twice as fast per query as with ranges of the caller's order).  A cloud that is already stored in scan order
    gets the same effect from shard_range."""
    order = spatial_order(points) if order is None else order
    lo, hi = shard_range(len(order), rank, world)
    return np.sort(order[lo:hi])


def shared_pair(m: int, config: int, regime: str, local_rank: int, world: int, barrier, tag: str = ""):
    """The seeded synthetic pair (synth.make_pair) for several ranks of one node: local rank 0 generates it once, the others
    map the same pages read-only (a 10^8-point pair is 4.8 GB; eight private copies plus the generator's temporaries are not
    needed).  `barrier` is a callable that synchronises the node's ranks."""
    import os
    from . import synth
    if world <= 1 or barrier is None:
        return synth.make_pair(m, config, regime)
    base = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"icpb_pair_{m}_{config}_{regime}_{os.environ.get('MASTER_PORT', '0')}{tag}")
    if local_rank == 0:
        src, tgt = synth.make_pair(m, config, regime)
        np.save(base + "_src.npy", src)
        np.save(base + "_tgt.npy", tgt)
    barrier()
    if local_rank != 0:
        src = np.load(base + "_src.npy", mmap_mode="r")
        tgt = np.load(base + "_tgt.npy", mmap_mode="r")
    barrier()
    if local_rank == 0:
        for suffix in ("_src.npy", "_tgt.npy"):  # (the mappings keep the pages alive)
            try:
                os.unlink(base + suffix)
            except OSError:
                pass
    return src, tgt


def exchange_unique_id(handle, dist, rank: int, src: int = 0) -> bytes:
    """Rank `src` creates the NCCL unique id through the C ABI (icp_comm_unique_id); everyone receives it."""
    box = [handle.comm_unique_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def init_sharded(handle, dist, rank: int, world: int):
    """icp_comm_init on every rank with rank 0's id."""
    uid = exchange_unique_id(handle, dist, rank)
    handle.comm_init(rank, world, uid)
