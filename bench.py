#!/usr/bin/env python
"""bench.py -- NN queries/s and ICP iterations/s of the B200-native ICP hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points M] [--regime primary] [--impl reference]

A *step* is one ICP iteration (exact NN for every source point -> 3-sigma rejection statistics -> Kabsch/SVD ->
apply) over the whole source cloud.  Workload at N=1: BASELINE.json configs[2], 10M <-> 10M points of the
synthetic "terrain+boxes" scene, source = target moved by the `primary` misalignment
(yaw 0.05 deg + 0.5 m) + 5 mm noise.  After W warm-up iterations the next K iterations of the same
registration are timed (device-resident inputs, CUDA events inside the library on its own stream, plus a
barrier + synchronize bracket; max over ranks).  `e2e` is the same loop through icp_register with HOST
buffers: H2D of both clouds, octree build, K iterations and the D2H write-back are all inside the timed region.
For N > 1 (torchrun) the source is sharded by point range, the target octree is replicated, and the two
per-iteration partial records are all-gathered with NCCL inside the library (strong scaling).

`--impl reference` times the reference's own CPU implementation (oracle/_ref when it was built in the build
container, else the oracle port) on the box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "nn_queries_per_s"
UNIT = "queries/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.gpu = gpu_index
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float) -> dict:
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # region shorter than the sampling period: fall back to every sample taken
            for ts, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------
def make_workload(m: int, regime: str):
    from iterativeclosestpoint_b200 import synth
    t0 = time.time()
    src, tgt = synth.make_pair(m, 3, regime)
    log(f"[bench] generated {m} <-> {m} points ({regime}) in {time.time() - t0:.1f}s")
    return src, tgt


def pinned_copy(a: np.ndarray):
    import torch
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    v = t.numpy()
    v[...] = a
    return t, v


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


NN_KERNELS = {
    0: "icpb::nn_kernel (literal traversal)", 1: "icpb::nn_kernel (climb)", 2: "icpb::nn_tile_kernel", 3: "icpb::nn_kernel (cell walk)",
    4: "icpb::nn_group_kernel + nn_kernel over its work list",
    5: "icpb::nn_keep_kernel + nn_collect_kernel + nn_kernel over their work lists",
    6: "icpb::nn_group_lean_kernel + nn_kernel over its work list while the registration moves; "
       "nn_keep_kernel + nn_collect_kernel once it has converged",
}


def ncu_traffic(m: int):
    """dram bytes per NN-kernel launch from the committed ncu --set full capture, if one exists for this size."""
    p = os.path.join(ROOT, "profiles", "nn_kernel_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get(str(m))
    return None


def cpu_reference_arm(src, tgt, sample: int, steps: int, warmup: int, budget_s: float = 25.0):
    """The reference's findNearest + the per-iteration statistics on a bounded sample, all host threads."""
    from oracle import binding
    kind = "reference" if binding.ref_available() else "port"
    orc = binding.Oracle()
    threads = orc.hw_threads()
    t0 = time.time()
    if kind == "reference":
        ref = binding.RefEngine()
        tree = ref.octree(tgt)
        find = lambda q: tree.find_nearest(q, nthreads=threads)
    else:
        tree = orc.octree(tgt)
        find = lambda q: tree.find_nearest(q, nthreads=threads)
    build_s = time.time() - t0
    sel = np.random.default_rng(1234).permutation(len(src))[:sample]
    cur = np.ascontiguousarray(src[sel])
    times = []
    total = warmup + steps
    spent = 0.0
    for it in range(total):
        t1 = time.time()
        idx = find(cur)
        dist, mask, st = orc.iteration_stats(cur, tgt, idx, it, 3.0, 0)
        a = cur[mask.astype(bool)]
        b = tgt[idx[mask.astype(bool)]]
        T = orc.kabsch(a, b)
        cur = orc.apply(T, cur)
        dt = time.time() - t1
        spent += dt
        if it >= warmup:
            times.append(dt)
        if spent > budget_s and len(times) >= 1:
            break
    ms = 1e3 * float(np.mean(times))
    return {"value": sample / (ms / 1e3), "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{sample} of {len(src)} source points against the full {len(tgt)}-point octree, "
                      f"{len(times)} iterations timed, octree build {build_s:.1f}s not included",
            "ms_per_step": ms, "steps_timed": len(times)}


# --------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--regime", default="primary")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--nn-mode", type=int, default=6)
    ap.add_argument("--shard", default="spatial", choices=["spatial", "blocks", "range"],
                    help="N > 1: point ranges of the spatially (Morton) ordered source, block-cyclic ranges of the caller's order "
                         "(64 Ki points per block), or one contiguous range of the caller's order per rank")
    ap.add_argument("--no-regimes", action="store_true", help="skip the near-converged and stress regimes (SURVEY.md 8(d))")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W, M = args.steps, max(args.warmup, 0), args.points
    workload = f"config3: {M}<->{M} pts terrain+boxes, {args.regime} misalignment, 50/1e-6/3sigma/leaf10/depth20"

    if args.impl == "reference":
        if rank != 0:
            return 0
        src, tgt = make_workload(M, args.regime)
        r = cpu_reference_arm(src, tgt, args.cpu_sample, K, W, budget_s=150.0)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "cpu_sample": args.cpu_sample},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist

    from iterativeclosestpoint_b200.engine import Handle, ICPParameters

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this implementation has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from iterativeclosestpoint_b200 import sharding

    src, tgt = make_workload(M, args.regime)
    shard_idx = None
    if args.shard == "spatial" and world > 1:
        shard_idx = sharding.shard_spatial(src, rank, world)  # host-side preparation, like generating the cloud
        ranges = None
        shard = np.ascontiguousarray(src[shard_idx])
    else:
        ranges = sharding.shard_blocks(M, rank, world) if (args.shard == "blocks" and world > 1) else [sharding.shard_range(M, rank, world)]
        shard = sharding.take_shard(src, ranges)

    h = Handle(local_rank)
    h.set_option("nn_mode", args.nn_mode)
    if world > 1:
        sharding.init_sharded(h, dist, rank, world)  # rank 0's NCCL id -> everyone -> icp_comm_init

    # ---- device-resident figure ---------------------------------------------------------------------------
    h.octree_build(tgt, 10, 20)
    info = h.octree_info()
    h.source_upload(shard)
    h.set_params(ICPParameters(maxIterations=max(W, 1), tolerance=0.0))
    barrier()
    if W > 0:
        h.register_resident(M)  # warm-up iterations (untimed); the source keeps its updated pose
    h.set_params(ICPParameters(maxIterations=K, tolerance=0.0))
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = h.kernel_launches()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    res = h.register_resident(M)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = h.kernel_launches() - launches0
    # device time of the timed region as the library's own events on its stream saw it
    loop_ms = float(res.timings_ms["loop"])
    nn_ms = float(res.timings_ms["nn_total"])
    iters = int(res.loopIterations)
    t = torch.tensor([loop_ms, nn_ms, float(iters)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    loop_ms, nn_ms, iters = float(t[0]), float(t[1]), int(t[2])
    clocks = sampler.summary(t_wall0, t_wall1)

    # ---- the other two regimes of SURVEY.md 8(d), same tree, same W + K resident iterations (reported, not the headline) ----
    regimes = {}
    if not args.no_regimes and args.regime == "primary" and M <= 20_000_000:
        from iterativeclosestpoint_b200 import synth
        for other in ("near", "stress"):
            rot, tr = synth.regime_transform(other)
            osrc = synth.make_source(tgt, synth.SEED_BASE + 3, rot, tr)
            h.source_upload(np.ascontiguousarray(osrc[shard_idx]) if shard_idx is not None else sharding.take_shard(osrc, ranges))
            del osrc
            if W > 0:
                h.set_params(ICPParameters(maxIterations=W, tolerance=0.0))
                barrier()
                h.register_resident(M)
            h.set_params(ICPParameters(maxIterations=K, tolerance=0.0))
            barrier()
            r3 = h.register_resident(M)
            barrier()
            t3 = torch.tensor([float(r3.timings_ms["loop"]), float(r3.timings_ms["nn_total"]), float(r3.loopIterations)],
                              dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            l3, n3, i3 = float(t3[0]), float(t3[1]), max(int(t3[2]), 1)
            regimes[other] = {"value": M * i3 / (l3 * 1e-3), "ms_per_step": l3 / i3, "nn_stage_ms": n3 / i3, "steps": i3}

    # ---- end to end through the C ABI with host buffers -------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        pin_s, host_src = pinned_copy(shard)
        pin_t, host_tgt = pinned_copy(tgt)
        h.set_params(ICPParameters(maxIterations=K, tolerance=0.0))
        barrier()
        t0 = time.perf_counter()
        if world > 1:
            r2 = h.register_sharded(host_src, M, host_tgt)
        else:
            r2 = h.register(host_src, host_tgt)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        it2 = max(int(r2.loopIterations), 1)
        e2e = {"value": M * it2 / dt, "unit": UNIT,
               "h2d_bytes_per_step": int((shard.nbytes + tgt.nbytes) * world / it2) if world == 1 else int(
                   (src.nbytes + tgt.nbytes * world) / it2),
               "d2h_bytes_per_step": int(src.nbytes / it2),
               "iterations": it2, "seconds": dt, "icp_iterations_per_s": it2 / dt,
               "breakdown_ms": {k: float(v) for k, v in r2.timings_ms.items()},
               "note": "icp_register/icp_register_sharded on pinned host buffers: H2D of source+target, octree build, "
                       f"{it2} iterations from the {args.regime} pose, D2H write-back; max over ranks"}
    sampler.stop()

    if rank == 0:
        peak, peak_src = peaks()
        n_local = len(shard)
        # algorithmic bytes of one NN-kernel launch on this rank: 24 B query read + 24 B
        # transformed query written back (apply fused into the load) + 4 B match + 8 B distance per query,
        # the 24 B/point target once, and the node table once.
        alg_bytes = 60 * n_local + 24 * M + int(info.node_bytes)
        nn_launch_ms = nn_ms / max(iters, 1)
        achieved = alg_bytes / (nn_launch_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": M * iters / (loop_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": iters,
            "warmup": W, "ms_per_step": loop_ms / max(iters, 1), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "l2": "inputs larger than L2 (source + sorted target + node table >> 126 MB)"
                       if M >= 4_000_000 else "inputs smaller than L2; no flush", "parallelism": f"source sharded x{world} ({'one point range per rank of the Morton-ordered source' if args.shard == 'spatial' and world > 1 else 'block-cyclic ranges of 65536 points' if args.shard == 'blocks' and world > 1 else 'one contiguous range per rank'}), octree replicated",
                       "octree": {"nodes": int(info.n_nodes), "leaves": int(info.n_leaves), "depth": int(info.depth),
                                  "build_ms": float(info.build_ms)},
                       "search": {"nodes": int(info.search_nodes), "depth": int(info.search_depth),
                                  "grid_levels": [int(info.grid_base_level), int(info.grid_fine_level)],
                                  "grid_base_cell_m": float(info.grid_base_cell), "grid_bytes": int(info.grid_bytes)},
                       "nn_mode": args.nn_mode},
            "icp_iterations_per_s": iters / (loop_ms * 1e-3),
            "nn_kernel_queries_per_s": M / (nn_launch_ms * 1e-3) if world == 1 else n_local * world / (nn_launch_ms * 1e-3),
            "nn_share_of_step": nn_ms / loop_ms,
            "clocks": clocks,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": NN_KERNELS.get(args.nn_mode, "icpb::nn_kernel"), "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(M), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": nn_launch_ms,
                         "note": "launch_ms is the whole NN stage of one iteration (the dominant kernel plus the kernels over its "
                                 "work lists), CUDA events inside the library on its stream; exact search is L1-wavefront / issue "
                                 "bound while the registration moves and HBM bound once it has converged (DESIGN.md 4)"},
        }
        if regimes:
            for k, v in regimes.items():
                if world == 1:  # (per rank the target term of the algorithmic bytes is not compulsory traffic)
                    v["roofline_frac"] = alg_bytes / (v["nn_stage_ms"] * 1e-3) / 1e9 / peak
            line["regimes"] = dict(regimes, note="same metric in the other two misalignment regimes of SURVEY.md 8(d): near = yaw 0.005 deg "
                                   "+ 5 cm (converged: the keep kernel settles every query, HBM bound), stress = yaw 5 deg + 0.5 m "
                                   "(edge points 43 m off: the climbing tree search carries the stage)")
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            log("[bench] timing the CPU reference arm on a bounded sample ...")
            r = cpu_reference_arm(src, tgt, args.cpu_sample, min(K, 3), 1, budget_s=25.0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    h.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
