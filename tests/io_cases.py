"""Seeded cases behind tests/golden/io_*.npz (the data-format rows of SURVEY.md 8(f)), shared by tools/make_golden_io.py
(which runs the compiled reference on them), the CPU tests (oracle vs golden) and the GPU tests (CUDA path vs golden)."""
import numpy as np

import clouds
from iterativeclosestpoint_b200 import synth


def _utm(n, seed):
    r = np.random.default_rng(seed)
    p = np.empty((n, 3))
    p[:, 0] = 431_250.125 + r.uniform(0, 800, n)
    p[:, 1] = 5_412_800.75 + r.uniform(0, 800, n)
    p[:, 2] = 212.0 + r.normal(0, 3, n)
    return p


def _signed(n, seed):
    r = np.random.default_rng(seed)
    p = r.uniform(-120.0, 90.0, (n, 3))
    p[::7] = np.round(p[::7], 3)          # values that sit on the 1 mm lattice: the truncating cast decides
    p[5] = [-0.0004, 0.0004, -0.9996]
    return p


def _edge(n, seed):
    """Out-of-range and non-finite coordinates: the writer's cast yields INT32_MIN (x86 cvttsd2si)."""
    p = _signed(n, seed)
    p[3] = [5.0e6, -5.0e6, 1.0]           # (p - min) / 0.001 > 2^31
    p[11, 2] = np.nan
    p[17, 1] = np.inf
    return p


# name -> (cloud maker, CLI scale, CLI offset)
IO_CLOUDS = {
    "terrain": (lambda: clouds.terrain(1500), (0.001, 0.001, 0.001), (-139.219, -137.327, 0.0)),   # test_icp.cpp:205-206 offsets
    "utm": (lambda: _utm(1200, 5), (0.01, 0.01, 0.001), (431_000.0, 5_412_000.0, 0.0)),
    "signed": (lambda: _signed(1000, 6), (0.001, 0.002, 0.0005), (-200.0, -150.5, -125.25)),
    "edge": (lambda: _edge(600, 7), (0.001, 0.001, 0.001), (0.0, 0.0, 0.0)),
}
DOWNSAMPLE_TARGETS = [1, 7, 333, 599, 600, 601, 5000]
STRIDES = [1, 3, 50, 10_000]
MAX_POINTS = [0, 1, 250, 10_000_000]


def transform_case(k):
    r = np.random.default_rng(100 + k)
    ang = r.uniform(-0.2, 0.2, 3)
    cx, cy, cz = np.cos(ang)
    sx, sy, sz = np.sin(ang)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = r.uniform(-3, 3, 3) * (1.0 if k else 1e5)
    return T


def foreign_las_image(n=700, record_length=34, offset_to_data=375, seed=9):
    """A LAS 1.2 file as other software writes it: point format 3 (34-byte records), VLR padding before the point block,
    signed raw coordinates -- what the readers must cope with beyond their own writers' output."""
    r = np.random.default_rng(seed)
    raw = r.integers(-2_000_000, 2_000_000, (n, 3), dtype=np.int64).astype(np.int32)
    img = np.zeros(offset_to_data + n * record_length, dtype=np.uint8)
    img[:4] = np.frombuffer(b"LASF", dtype=np.uint8)
    img[24], img[25] = 1, 2
    img[94:96] = np.frombuffer(np.uint16(227).tobytes(), dtype=np.uint8)
    img[96:100] = np.frombuffer(np.uint32(offset_to_data).tobytes(), dtype=np.uint8)
    img[104] = 3
    img[105:107] = np.frombuffer(np.uint16(record_length).tobytes(), dtype=np.uint8)
    img[107:111] = np.frombuffer(np.uint32(n).tobytes(), dtype=np.uint8)
    scale = np.array([0.01, 0.01, 0.001]); offset = np.array([500_000.0, 4_100_000.0, -12.5])
    img[131:155] = np.frombuffer(scale.tobytes(), dtype=np.uint8)
    img[155:179] = np.frombuffer(offset.tobytes(), dtype=np.uint8)
    img[227:offset_to_data] = r.integers(0, 256, offset_to_data - 227, dtype=np.uint8)
    body = r.integers(0, 256, (n, record_length), dtype=np.uint8)
    body[:, :12] = raw.view(np.uint8).reshape(n, 12)
    img[offset_to_data:] = body.reshape(-1)
    return img
