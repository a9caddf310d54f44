"""Numpy restatements of what the device code does with the gathered per-iteration records (iter.cu: stat_merge /
stat_a_finalize / solve_step): ranks are merged IN RANK ORDER with the same arithmetic on every rank, so all ranks derive
bit-identical statistics and therefore the same transform without a broadcast.  Test support for tests/test_sharding_gloo.py
(world_size-2 gloo runs on CPU prove the host plumbing; the CUDA path itself is checked by tools/sharded_check.py on GPUs)."""
import numpy as np


# ---- stage A: Chan partials (n, mean, M2, min, max, problems) -----------------------------------------------------------
def stat_partial(d: np.ndarray) -> np.ndarray:
    """One rank's partial over its distances (two-pass form of what stat_a_kernel produces)."""
    d = np.asarray(d, dtype=np.float64)
    fin = np.isfinite(d)
    n = float(d.size)
    mean = float(d.mean()) if d.size else 0.0
    m2 = float(((d - mean) ** 2).sum()) if d.size else 0.0
    dmin = float(d[fin].min()) if fin.any() else np.finfo(np.float64).max
    dmax = float(d[fin].max()) if fin.any() else 0.0
    return np.array([n, mean, m2, dmin, dmax, float((~fin).sum())])


def stat_merge(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """common.cuh: stat_merge."""
    n = a[0] + b[0]
    if n == 0.0:
        mean, m2 = 0.0, 0.0
    else:
        delta = b[1] - a[1]
        f = b[0] / n
        mean = a[1] + delta * f
        m2 = a[2] + b[2] + delta * delta * a[0] * f
    return np.array([n, mean, m2, min(a[3], b[3]), max(a[4], b[4]), a[5] + b[5]])


def merge_in_rank_order(parts) -> np.ndarray:
    acc = np.asarray(parts[0], dtype=np.float64)
    for p in parts[1:]:
        acc = stat_merge(acc, np.asarray(p, dtype=np.float64))
    return acc


def threshold(merged: np.ndarray, n_global: int, sigma: float, iteration: int, engine_variant: bool = True):
    """mean, population std, rejection threshold (icpengine.cpp:235-255; CLI icp_registration.cpp:510-523)."""
    mean = merged[1]
    std = float(np.sqrt(merged[2] / float(n_global)))
    if engine_variant and iteration == 0:
        thr = mean + max(sigma * std, 0.5 * mean)
    else:
        thr = mean + sigma * std
    return mean, std, thr


# ---- stage B: 17 doubles (count, sum d^2, sum(a-p)[3], sum(b-q)[3], sum (a-p)(b-q)^T [9]) ------------------------------------
def moment_partial(a: np.ndarray, b: np.ndarray, d: np.ndarray, thr: float, pa: np.ndarray, pb: np.ndarray) -> np.ndarray:
    ok = d <= thr
    aa = a[ok] - pa
    bb = b[ok] - pb
    out = np.zeros(17)
    out[0] = float(ok.sum())
    out[1] = float((d[ok] * d[ok]).sum())
    out[2:5] = aa.sum(0)
    out[5:8] = bb.sum(0)
    out[8:17] = (aa.T @ bb).reshape(9)
    return out


def sum_in_rank_order(parts) -> np.ndarray:
    acc = np.array(parts[0], dtype=np.float64)
    for p in parts[1:]:
        acc = acc + np.asarray(p, dtype=np.float64)
    return acc


def moments_to_H(m: np.ndarray, pa: np.ndarray, pb: np.ndarray):
    """centroids and cross-covariance from pivoted sums (iter.cu: moments_to_H)."""
    n = m[0]
    ma, mb = m[2:5] / n, m[5:8] / n
    H = m[8:17].reshape(3, 3) - np.outer(m[2:5], mb)
    return pa + ma, pb + mb, H
