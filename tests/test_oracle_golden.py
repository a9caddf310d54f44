"""CPU tests: the oracle (oracle/icp_oracle.c, this repo's restatement of the reference's hot path) against the
golden vectors under tests/golden/, which tools/make_golden.py produced by running the UNMODIFIED compiled
reference.  This is what pins the oracle (the reference itself ships no tests / golden vectors).

Bars: octree structure, boxes, NN indices, SVD given H, apply, 4x4 product: BIT-EXACT.  Anything downstream of the
3xN.Nx3 cross-covariance (which the reference sends through Eigen's blocked GEMM) agrees to 1e-12 relative;
iteration counts, inlier counts and exit paths are identical.
"""
import os

import numpy as np
import pytest

import clouds
from golden_cases import CLI_RUNS, ENGINE_RUNS, TREE_CASES, kabsch_inputs, svd_inputs
from iterativeclosestpoint_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def dig(*a):
    return np.frombuffer(bytes.fromhex(synth.digest(*a)), dtype=np.uint8)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, float(np.max(np.abs(b)))))


@pytest.mark.parametrize("name,make,leaf,depth", TREE_CASES, ids=[c[0] for c in TREE_CASES])
def test_oracle_octree_and_nn_match_reference_vectors(oracle, name, make, leaf, depth):
    g = load("tree_" + name)
    tgt = make()
    assert np.array_equal(g["digest"], dig(tgt)), "input generator drifted: regenerate with tools/make_golden.py"
    tree = oracle.octree(tgt, leaf, depth)
    d = tree.dump()
    for k in ("depth", "key", "leaf", "count", "idx", "box"):
        assert np.array_equal(d[k], g["tree_" + k]), f"{name}: tree field {k}"
    for qname, q in clouds.query_sets(tgt).items():
        q = q[:1500]
        assert np.array_equal(tree.find_nearest(q), g["nn_engine_" + qname]), f"{name}/{qname}"
        if "nn_cli_" + qname in g.files:
            assert np.array_equal(tree.find_nearest(q, variant=1), g["nn_cli_" + qname]), f"{name}/{qname} (CLI)"
    if name == "lattice_exact":
        q = clouds.lattice_tie_queries(tgt)
        assert np.array_equal(tree.find_nearest(q), g["nn_engine_ties"])
        assert np.array_equal(tree.find_nearest(q, variant=1), g["nn_cli_ties"])
        # the fixture really exercises cross-leaf ties that the lowest index does NOT win
        d2 = ((q[:, None, :] - tgt[None, :, :]) ** 2).sum(-1)
        tied = (d2 == d2.min(1, keepdims=True)).sum(1) > 1
        assert tied.sum() > 200 and (g["nn_engine_ties"][tied] != d2.argmin(1)[tied]).sum() > 0
    tree.close()


@pytest.mark.parametrize("name,make,kw", ENGINE_RUNS, ids=[c[0] for c in ENGINE_RUNS])
def test_oracle_engine_runs_match_reference_vectors(oracle, name, make, kw):
    g = load("engine_" + name)
    src, tgt = make()
    assert np.array_equal(g["digest"], dig(src, tgt))
    r = oracle.icp(src, tgt, **kw)
    assert r.status == int(g["status"]) and r.success == bool(g["success"])
    assert r.total_iterations == int(g["total_iterations"])
    n_hist = len(g["hist_rmse"])
    assert len(r.history) == n_hist
    for k, h in enumerate(r.history):
        assert h.iteration == g["hist_iteration"][k]
        assert h.valid_points == g["hist_valid"][k] and h.outlier_points == g["hist_outlier"][k]
        assert h.has_angles == bool(g["hist_has_angles"][k])
        assert abs(h.rmse - g["hist_rmse"][k]) <= 1e-12 * max(g["hist_rmse"][k], 1e-300)
        assert rel(h.transform, g["hist_T"][k]) <= 1e-12
        if h.has_angles and g["hist_angle"][k] > 1e-3:
            # acos near 1 amplifies the 1e-16-level difference of the trace by 1/sin(angle)
            ga = float(g["hist_angle"][k])
            assert abs(h.rotation_angle - ga) <= 1e-9 * ga + 1e-12 / np.radians(ga)
            assert abs(h.translation_distance - g["hist_trans"][k]) <= 1e-12 * max(1.0, g["hist_trans"][k])
    assert abs(r.final_rmse - float(g["final_rmse"])) <= 1e-12 * max(float(g["final_rmse"]), 1e-300)
    if r.success:
        assert rel(r.final_R, g["final_R"]) <= 1e-12
        assert np.max(np.abs(r.final_t - g["final_t"])) <= 1e-12 * max(1.0, float(np.max(np.abs(g["final_t"]))))
        s = r.source_out[::37]
        assert np.max(np.abs(s - g["source_out_sample"])) <= 1e-12 * float(np.max(np.abs(s)))
    else:
        # failure exits leave the source untouched (icpengine.cpp:160-164, 319-323)
        assert np.array_equal(r.source_out, src)
        assert np.array_equal(g["source_out_digest"], dig(src))


def test_first_iteration_is_bit_exact_with_reference(oracle):
    """Before any Eigen GEMM result feeds back, every loop quantity is a plain sequential loop: iteration 1's RMSE
    (sequential sum over inliers) and inlier counts must equal the reference's to the last bit."""
    for name, make, kw in ENGINE_RUNS[:4]:
        g = load("engine_" + name)
        src, tgt = make()
        r = oracle.icp(src, tgt, **kw)
        assert r.history[0].rmse == g["hist_rmse"][0], name
        assert r.history[0].valid_points == g["hist_valid"][0]


@pytest.mark.parametrize("name,make,kw", CLI_RUNS, ids=[c[0] for c in CLI_RUNS])
def test_oracle_cli_runs_match_reference_vectors(oracle, name, make, kw):
    g = load("cli_" + name)
    src, tgt = make()
    assert np.array_equal(g["digest"], dig(src, tgt))
    r = oracle.icp(src, tgt, variant=1, **kw)
    assert len(r.history) == len(g["iteration_T"])
    for h, T in zip(r.history, g["iteration_T"]):
        assert rel(h.transform, T) <= 1e-12
    assert rel(r.final_R, g["final_R"]) <= 1e-11   # CLI returns the LAST incremental transform (:616-621)
    assert np.max(np.abs(r.final_t - g["final_t"])) <= 1e-11
    s = r.source_out[::37]
    assert np.max(np.abs(s - g["source_out_sample"])) <= 1e-12 * float(np.max(np.abs(s)))


def test_oracle_svd_kabsch_apply_match_reference_vectors(oracle):
    g = load("kabsch_svd")
    Hs, cAs, cBs = svd_inputs()
    assert np.array_equal(g["digest"], dig(Hs, cAs, cBs))
    for i in range(len(Hs)):
        U, S, V = oracle.svd3(Hs[i])
        assert np.array_equal(U, g["U"][i]) and np.array_equal(S, g["S"][i]) and np.array_equal(V, g["V"][i]), i
        assert np.array_equal(oracle.solve_from_H(Hs[i], cAs[i], cBs[i]), g["T"][i]), i
    for i, (a, b) in enumerate(kabsch_inputs()):
        assert np.array_equal(g[f"kabsch{i}_digest"], dig(a, b))
        cA, cB, H = oracle.centroids_H(a, b)
        assert np.array_equal(cA, g[f"kabsch{i}_cA"]) and np.array_equal(cB, g[f"kabsch{i}_cB"])  # sequential means
        assert rel(H, g[f"kabsch{i}_H"]) <= 1e-12                                                # Eigen GEMM order
        T = oracle.kabsch(a, b)
        if i < 3:  # full-rank clouds; the planar one has a free sign in the null direction
            assert rel(T, g[f"kabsch{i}_T_engine"]) <= 1e-9
            assert rel(T, g[f"kabsch{i}_T_cli"]) <= 1e-9
        Tm = g[f"kabsch{i}_T_engine"]
        assert np.array_equal(oracle.apply(Tm, a)[::11], g[f"kabsch{i}_applied_sample"])
        assert np.array_equal(oracle.mat4_mul(Tm, Tm), g[f"kabsch{i}_TT"])
        ang, tr = oracle.angles(Tm)
        assert tr == g[f"kabsch{i}_angles"][1]
        assert abs(ang - g[f"kabsch{i}_angles"][0]) <= 1e-9 * max(1.0, abs(g[f"kabsch{i}_angles"][0]))
