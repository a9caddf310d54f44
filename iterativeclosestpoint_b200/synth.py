"""Seeded synthetic clouds for parity tests and bench.py ("Synthetic inputs").

Host-side numpy only; the arrays produced here are handed bit-identically to the CUDA path, the oracle
and the compiled reference.  RNG = splitmix64 in counter mode -> U[0,1) doubles -> Box-Muller normals.

Scene "terrain+boxes": tile side E = sqrt(M / rho) with rho = 10 pts/m^2; 80 % of the points lie on the
2.5-D terrain z = 3 sin(0.1 x) + 2 cos(0.07 y) + 0.5 sin(0.5 x + 0.3 y), 20 % on the walls/roofs of
ceil(M / 50 000) axis-aligned 10 x 10 x 5 m boxes.  The source is the target rotated about the tile
centre, translated, plus N(0, (5 mm)^2) per coordinate; `regime` picks the misalignment.
"""
from __future__ import annotations

import hashlib
import math

import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)

SEED_BASE = 20260000
DENSITY = 10.0  # points per square metre
# LAS offsets used by the reference's data generator (test_icp.cpp:205-206).
LAS_OFFSET = (-139.219, -137.327, 0.0)
LAS_SCALE = 0.001

# (yaw degrees, translation metres along (0.6,-0.6,0.5)/|.|) per regime, SURVEY.md 8(d)
REGIMES = {
    "stress": (5.0, 0.5),
    "primary": (0.05, 0.5),
    "near": (0.005, 0.05),
    "converged": (0.0, 0.0),
}


def splitmix64(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """n consecutive splitmix64 outputs of the stream `seed`, starting at draw number `offset`."""
    with np.errstate(over="ignore"):
        k = np.arange(offset + 1, offset + n + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + k * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


class Stream:
    """Sequential consumer of one splitmix64 stream."""

    def __init__(self, seed: int):
        self.seed = seed & 0xFFFFFFFFFFFFFFFF
        self.pos = 0

    def bits(self, n: int) -> np.ndarray:
        out = splitmix64(self.seed, n, self.pos)
        self.pos += n
        return out

    def uniform(self, n: int) -> np.ndarray:
        return (self.bits(n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    def normal(self, n: int) -> np.ndarray:
        m = (n + 1) // 2
        u1 = 1.0 - self.uniform(m)  # (0,1]
        u2 = self.uniform(m)
        r = np.sqrt(-2.0 * np.log(u1))
        out = np.empty(2 * m, dtype=np.float64)
        out[0::2] = r * np.cos(2.0 * math.pi * u2)
        out[1::2] = r * np.sin(2.0 * math.pi * u2)
        return out[:n]


def terrain_z(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    return 3.0 * np.sin(0.1 * x) + 2.0 * np.cos(0.07 * y) + 0.5 * np.sin(0.5 * x + 0.3 * y)


def tile_side(m: int) -> float:
    return math.sqrt(m / DENSITY)


def make_target(m: int, seed: int, las_quantise: bool = False, tile: float | None = None) -> np.ndarray:
    """(m,3) float64 C-contiguous target cloud of the terrain+boxes scene."""
    rs = Stream(seed)
    e = tile_side(m) if tile is None else float(tile)
    n_box_pts = m // 5
    n_ter = m - n_box_pts
    pts = np.empty((m, 3), dtype=np.float64)
    x = rs.uniform(n_ter) * e
    y = rs.uniform(n_ter) * e
    pts[:n_ter, 0] = x
    pts[:n_ter, 1] = y
    pts[:n_ter, 2] = terrain_z(x, y)
    if n_box_pts:
        n_boxes = max(1, -(-m // 50000))
        span = max(e - 10.0, 0.0)
        bx = rs.uniform(n_boxes) * span
        by = rs.uniform(n_boxes) * span
        bz = terrain_z(bx + 5.0, by + 5.0)
        which = np.minimum((rs.uniform(n_box_pts) * n_boxes).astype(np.int64), n_boxes - 1)
        face = rs.uniform(n_box_pts) * 300.0  # 4 walls of 50 m^2 + a 100 m^2 roof
        u = rs.uniform(n_box_pts)
        v = rs.uniform(n_box_pts)
        px = np.empty(n_box_pts)
        py = np.empty(n_box_pts)
        pz = np.empty(n_box_pts)
        w0 = face < 50.0
        w1 = (face >= 50.0) & (face < 100.0)
        w2 = (face >= 100.0) & (face < 150.0)
        w3 = (face >= 150.0) & (face < 200.0)
        rf = face >= 200.0
        px[w0] = 0.0; py[w0] = 10.0 * u[w0]; pz[w0] = 5.0 * v[w0]
        px[w1] = 10.0; py[w1] = 10.0 * u[w1]; pz[w1] = 5.0 * v[w1]
        px[w2] = 10.0 * u[w2]; py[w2] = 0.0; pz[w2] = 5.0 * v[w2]
        px[w3] = 10.0 * u[w3]; py[w3] = 10.0; pz[w3] = 5.0 * v[w3]
        px[rf] = 10.0 * u[rf]; py[rf] = 10.0 * v[rf]; pz[rf] = 5.0
        pts[n_ter:, 0] = bx[which] + px
        pts[n_ter:, 1] = by[which] + py
        pts[n_ter:, 2] = bz[which] + pz
        # interleave the box points among the terrain points so index order carries no structure
        perm = np.argsort(rs.bits(m), kind="stable")
        pts = pts[perm]
    if las_quantise:
        pts = las_truncate(pts)
    return np.ascontiguousarray(pts)


def las_truncate(pts: np.ndarray, offset=LAS_OFFSET, scale: float = LAS_SCALE) -> np.ndarray:
    """The LAS writer's quantisation: int32((v - offset) / scale) by truncation (test_icp.cpp:139-141)."""
    off = np.asarray(offset, dtype=np.float64)
    q = np.trunc((pts - off) / scale).astype(np.int32)
    return q.astype(np.float64) * scale + off


def rotation_zyx(yaw: float, pitch: float, roll: float) -> np.ndarray:
    cy, sy = math.cos(yaw), math.sin(yaw)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cr, sr = math.cos(roll), math.sin(roll)
    rz = np.array([[cy, -sy, 0.0], [sy, cy, 0.0], [0.0, 0.0, 1.0]])
    ry = np.array([[cp, 0.0, sp], [0.0, 1.0, 0.0], [-sp, 0.0, cp]])
    rx = np.array([[1.0, 0.0, 0.0], [0.0, cr, -sr], [0.0, sr, cr]])
    return rz @ ry @ rx


def make_source(target: np.ndarray, seed: int, rot: np.ndarray, trans, noise_sigma: float = 0.005,
                centre=None, las_quantise: bool = False) -> np.ndarray:
    """Rigidly moved + noised copy of `target` (rotation about `centre`, default the bbox centre)."""
    rs = Stream(seed ^ 0x9E3779B97F4A7C15)
    c = (target.min(axis=0) + target.max(axis=0)) * 0.5 if centre is None else np.asarray(centre, dtype=np.float64)
    src = (target - c) @ np.asarray(rot, dtype=np.float64).T + c + np.asarray(trans, dtype=np.float64)
    if noise_sigma > 0.0:
        src = src + noise_sigma * rs.normal(3 * len(target)).reshape(-1, 3)
    if las_quantise:
        src = las_truncate(src)
    return np.ascontiguousarray(src)


def regime_transform(regime: str):
    yaw_deg, tr = REGIMES[regime]
    d = np.array([0.6, -0.6, 0.5])
    d = d / np.linalg.norm(d)
    return rotation_zyx(math.radians(yaw_deg), 0.0, 0.0), tr * d


def make_pair(m: int, config: int = 2, regime: str = "stress", n_src: int | None = None,
              las_quantise: bool = False, noise_sigma: float = 0.005, seed: int | None = None):
    """(source, target) for BASELINE.json config numbers 2-4 at `m` target points."""
    seed = SEED_BASE + config if seed is None else seed
    tgt = make_target(m, seed, las_quantise=las_quantise)
    rot, tr = regime_transform(regime)
    src = make_source(tgt, seed, rot, tr, noise_sigma=noise_sigma, las_quantise=las_quantise)
    if n_src is not None and n_src < len(src):
        src = np.ascontiguousarray(src[:n_src])
    return src, tgt


def make_test_icp_pair(m: int = 10000, seed: int = SEED_BASE + 1):
    """BASELINE.json config #1: the recipe of test_icp.cpp:165-189,211-215 with a fixed seed instead of
    time(NULL): yaw U[0,10 deg], pitch/roll +-yaw/2, t = (+-2.5, +-2.5, +-1) m, both clouds LAS-truncated."""
    tgt = make_target(m, seed, las_quantise=True)
    rs = Stream(seed ^ 0xD1B54A32D192ED03)
    u = rs.uniform(6)
    yaw = math.radians(10.0 * u[0])
    pitch = (u[1] - 0.5) * yaw
    roll = (u[2] - 0.5) * yaw
    t = np.array([(u[3] - 0.5) * 5.0, (u[4] - 0.5) * 5.0, (u[5] - 0.5) * 2.0])
    rot = rotation_zyx(yaw, pitch, roll)
    src = make_source(tgt, seed, rot, t, noise_sigma=0.005, las_quantise=True)
    return src, tgt


def make_lattice(side: int = 24, pitch: float = 0.025, seed: int = 7) -> np.ndarray:
    """LAS-quantised lattice with heavy exact distance ties (adversarial tie-break set)."""
    g = np.arange(side, dtype=np.float64) * pitch
    x, y = np.meshgrid(g, g, indexing="ij")
    rs = Stream(seed)
    z = np.round(rs.uniform(side * side).reshape(side, side) * 4.0) * pitch
    pts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    perm = np.argsort(rs.bits(len(pts)), kind="stable")
    return np.ascontiguousarray(las_truncate(pts[perm] + 1000.0, offset=(0.0, 0.0, 0.0)))


def small_pair(pair: int, n: int = 2000, seed: int = SEED_BASE + 5):
    """Pair `pair` of config #5 (4096 independent 2k-point registrations on 14 m tiles)."""
    s = (seed * 4096 + pair) & 0xFFFFFFFFFFFFFFFF
    tgt = make_target(n, s, tile=14.0)
    rs = Stream(s ^ 0xA0761D6478BD642F)
    u = rs.uniform(4)
    rot = rotation_zyx(math.radians(3.0 * u[0]), 0.0, 0.0)
    t = np.array([(u[1] - 0.5) * 0.6, (u[2] - 0.5) * 0.6, (u[3] - 0.5) * 0.3])
    src = make_source(tgt, s, rot, t)
    return src, tgt


def digest(*arrays: np.ndarray) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]
