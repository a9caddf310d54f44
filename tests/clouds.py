"""Seeded inputs shared by the CPU and GPU parity tests (sizes the oracle finishes in seconds)."""
import numpy as np

from iterativeclosestpoint_b200 import synth


def rng(seed):
    return np.random.default_rng(seed)


def terrain(m=20000, seed=11, las=False):
    return synth.make_target(m, seed, las_quantise=las)


def duplicates(m=6000, seed=12):
    """25 % of the points are exact copies of other points (ties inside leaves)."""
    base = synth.make_target(m, seed)
    r = rng(seed)
    dup = r.integers(0, m, size=m // 4)
    pos = r.integers(0, m, size=m // 4)
    base[pos] = base[dup]
    return np.ascontiguousarray(base)


def coincident(m=500):
    """All points identical: depth-capped leaves holding more than max_pts points."""
    return np.ascontiguousarray(np.tile(np.array([[3.25, -1.5, 0.75]]), (m, 1)))


def two_clusters(m=3000, seed=13):
    """Two tight clusters far apart: long single-child chains down to max_depth."""
    r = rng(seed)
    a = r.normal(size=(m // 2, 3)) * 1e-7
    b = r.normal(size=(m - m // 2, 3)) * 1e-7 + 1000.0
    return np.ascontiguousarray(np.concatenate([a, b])[r.permutation(m)])


def planar(m=4000, seed=14):
    r = rng(seed)
    p = r.uniform(0, 20, size=(m, 3))
    p[:, 2] = 1.0
    return np.ascontiguousarray(p)


def collinear(m=300, seed=15):
    r = rng(seed)
    t = r.uniform(0, 50, size=m)
    return np.ascontiguousarray(np.stack([t, 2.0 * t + 1.0, -0.5 * t], axis=1))


def lattice():
    return synth.make_lattice(side=24, pitch=0.025, seed=7)


def lattice_exact(side=20, seed=17):
    """Lattice whose coordinates are dyadic rationals (pitch 0.25), so squared distances to the tie queries
    below are EXACTLY equal in floating point; index order is shuffled."""
    r = rng(seed)
    g = np.arange(side, dtype=np.float64) * 0.25 + 64.0
    x, y = np.meshgrid(g, g, indexing="ij")
    z = r.integers(0, 3, size=side * side).astype(np.float64) * 0.25 + 8.0
    pts = np.stack([x.ravel(), y.ravel(), z], axis=1)
    return np.ascontiguousarray(pts[r.permutation(len(pts))])


def lattice_tie_queries(lat, seed=16, n=3000):
    """Queries on lattice nodes / edge midpoints / cell centres (offsets are multiples of half the pitch):
    many exactly equidistant targets."""
    r = rng(seed)
    base = lat[r.integers(0, len(lat), size=n)]
    off = r.integers(0, 2, size=(n, 3)) * 0.125
    return np.ascontiguousarray(base + off)


def query_sets(tgt, seed=21):
    """Named query clouds around a target: near-converged, offset, far above, far outside, exact copies."""
    r = rng(seed)
    m = len(tgt)
    sel = r.permutation(m)[: min(m, 6000)]
    out = {}
    out["copies"] = tgt[sel].copy()
    out["noise5mm"] = tgt[sel] + r.normal(size=(len(sel), 3)) * 0.005
    rot, tr = synth.regime_transform("stress")
    c = (tgt.min(0) + tgt.max(0)) * 0.5
    out["stress"] = (tgt[sel] - c) @ rot.T + c + tr
    out["above"] = tgt[sel] + np.array([0.0, 0.0, 40.0])
    out["outside"] = tgt[sel] * 3.0 + np.array([500.0, -300.0, 20.0])
    out["uniform"] = r.uniform(tgt.min(0) - 5.0, tgt.max(0) + 5.0, size=(len(sel), 3))
    return {k: np.ascontiguousarray(v) for k, v in out.items()}
