"""The seeded cases behind tests/golden/*.npz, shared by tools/make_golden.py (which runs the compiled
reference on them), the CPU tests (oracle vs golden) and the GPU tests (CUDA path vs golden)."""
import numpy as np

import clouds
from iterativeclosestpoint_b200 import synth

# (name, target maker, octreeMaxPoints, octreeMaxDepth)
TREE_CASES = [
    ("terrain", lambda: clouds.terrain(6000), 10, 20),
    ("terrain_las", lambda: clouds.terrain(4000, las=True), 10, 20),
    ("leaf5_depth6", lambda: clouds.terrain(4000), 5, 6),
    ("duplicates", lambda: clouds.duplicates(3000), 10, 20),
    ("coincident", lambda: clouds.coincident(200), 10, 20),
    ("two_clusters", lambda: clouds.two_clusters(1500), 10, 20),
    ("lattice_exact", clouds.lattice_exact, 10, 20),
]


def _tiny():
    return (np.array([[0.0, 0, 0], [1.0, 0, 0]]), np.array([[0, 0, 0.1], [1, 0, 0.1], [0, 1, 0.1]], dtype=np.float64))


# (name, (source, target) maker, kwargs of the engine run)
ENGINE_RUNS = [
    ("config1_3k", lambda: synth.make_test_icp_pair(3000), dict(max_iterations=50, tolerance=1e-6, sigma=3.0)),
    ("near_4k", lambda: synth.make_pair(4000, 2, "near"), dict(max_iterations=30)),
    ("stress_3k_maxiter4", lambda: synth.make_pair(3000, 2, "stress"), dict(max_iterations=4)),
    ("primary_3k_sigma2_leaf5", lambda: synth.make_pair(3000, 2, "primary"),
     dict(max_iterations=12, sigma=2.0, leaf=5, depth=12)),
    ("too_few_inliers", _tiny, dict()),
    ("cancel_after_2", lambda: synth.make_pair(2000, 2, "near"), dict(max_iterations=10, stop_after=2)),
]

CLI_RUNS = [
    ("config1_3k", lambda: synth.make_test_icp_pair(3000, seed=77), dict(max_iterations=20, tolerance=1e-2)),
    ("near_2k", lambda: synth.make_pair(2000, 2, "near"), dict(max_iterations=8, tolerance=1e-5)),
]


def svd_inputs(n=120, seed=5):
    r = np.random.default_rng(seed)
    Hs = np.empty((n, 3, 3)); cAs = np.empty((n, 3)); cBs = np.empty((n, 3))
    for i in range(n):
        H = r.normal(size=(3, 3)) * 10 ** r.uniform(-3, 6)
        if i % 7 == 0:
            H[:, 2] = H[:, 0] * 2      # rank deficient
        if i % 11 == 0:
            H = np.diag(r.normal(size=3))
        if i % 13 == 0:
            H = -np.abs(H)             # reflection branch
        if i == 0:
            H = np.zeros((3, 3))
        Hs[i] = H
        cAs[i] = r.normal(size=3) * 100
        cBs[i] = r.normal(size=3) * 100
    return Hs, cAs, cBs


def kabsch_inputs(seed=6):
    r = np.random.default_rng(seed)
    R = synth.rotation_zyx(0.3, -0.1, 0.2)
    out = []
    for n, offset in ((3, 0.0), (50, 10.0), (5000, 5e5)):
        a = r.normal(size=(n, 3)) * 30 + offset
        b = a @ R.T + np.array([1.0, -2.0, 0.5]) + r.normal(size=(n, 3)) * 0.01
        out.append((np.ascontiguousarray(a), np.ascontiguousarray(b)))
    pl = clouds.planar(1500)
    out.append((pl, np.ascontiguousarray(pl @ R.T + 1.0)))
    return out
