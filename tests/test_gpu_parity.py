"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.

Bars: tree structure, NN indices, distances, inlier masks, SVD given H and transform apply
are BIT-EXACT; statistics that involve a sum over N agree to 1e-12 relative (the reference sums sequentially,
the GPU in a fixed tree order); end-to-end runs have identical iteration counts and T within 1e-9.
"""
import ctypes as C

import numpy as np
import pytest

import clouds
from iterativeclosestpoint_b200 import synth
from iterativeclosestpoint_b200.engine import (ICP, Handle, ICPEngine, ICPParameters, Octree, VARIANT_CLI,
                                               VARIANT_ENGINE, best_fit_transform)

pytestmark = pytest.mark.gpu

REL_SUM = 1e-12   # tolerance for quantities that contain a length-N floating-point sum
REL_E2E = 1e-9    # north_star: final transform within 1e-9 relative


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, float(np.max(np.abs(b)))))


TREE_CASES = [
    ("terrain20k", lambda: clouds.terrain(20000), 10, 20),
    ("terrain_las", lambda: clouds.terrain(8000, las=True), 10, 20),
    ("terrain_leaf5_depth6", lambda: clouds.terrain(8000), 5, 6),
    ("terrain_leaf100", lambda: clouds.terrain(8000), 100, 21),
    ("duplicates", clouds.duplicates, 10, 20),
    ("coincident", clouds.coincident, 10, 20),
    ("two_clusters", clouds.two_clusters, 10, 20),
    ("planar", clouds.planar, 10, 20),
    ("collinear", clouds.collinear, 10, 20),
    ("lattice", clouds.lattice, 10, 20),
    ("lattice_exact", clouds.lattice_exact, 10, 20),
    ("single", lambda: np.array([[1.0, 2.0, 3.0]]), 10, 20),
    ("eleven", lambda: clouds.terrain(11), 10, 20),
    ("depth0", lambda: clouds.terrain(500), 10, 0),
    # deeper than one 64-bit key word (the reference's GUI offers octreeMaxDepth 10 .. 50, settingspage.cpp:76): clusters of
    # coincident points are split all the way down
    ("coincident_depth30", clouds.coincident, 10, 30),
    ("two_clusters_depth30", clouds.two_clusters, 10, 30),
    ("duplicates_leaf1_depth50", clouds.duplicates, 1, 50),
    ("terrain_depth45", lambda: clouds.terrain(8000), 10, 45),
]


@pytest.mark.parametrize("name,make,leaf,depth", TREE_CASES, ids=[c[0] for c in TREE_CASES])
def test_octree_structure_bit_exact(handle, oracle, name, make, leaf, depth):
    tgt = make()
    handle.octree_build(tgt, leaf, depth)
    got = handle.octree_dump()
    want = oracle.octree(tgt, leaf, depth).dump()
    for k in ("depth", "key", "leaf", "count", "idx"):
        assert np.array_equal(got[k], want[k]), f"{name}: node table field {k} differs"
    assert np.array_equal(got["box"], want["box"]), f"{name}: boxes differ (must be bit-identical bisections)"
    info = handle.octree_info()
    assert info.n_nodes == len(want["depth"]) and info.n_leaves == int(want["leaf"].sum())
    assert info.depth == int(want["depth"].max())


@pytest.mark.parametrize("mode", [0, 3, 4, 5, 6], ids=["literal", "walk", "group", "keep", "auto"])
@pytest.mark.parametrize("name,make,leaf,depth", TREE_CASES, ids=[c[0] for c in TREE_CASES])
def test_nn_indices_bit_exact(handle, oracle, name, make, leaf, depth, mode):
    tgt = make()
    handle.set_option("nn_mode", mode)
    handle.octree_build(tgt, leaf, depth)
    otree = oracle.octree(tgt, leaf, depth)
    for qname, q in clouds.query_sets(tgt).items():
        idx, dist, _ = handle.nn_query(q)
        want = otree.find_nearest(q)
        bad = np.flatnonzero(idx != want)
        assert bad.size == 0, f"{name}/{qname}/mode{mode}: {bad.size} of {len(q)} indices differ, first {bad[:5]}"
        # computeDistance: sqrt(dx*dx+dy*dy+dz*dz) -- numpy evaluates the same expression in the same order
        dx = q - tgt[want]
        d = np.sqrt(dx[:, 0] * dx[:, 0] + dx[:, 1] * dx[:, 1] + dx[:, 2] * dx[:, 2])
        assert np.array_equal(dist, d), f"{name}/{qname}: distances are not bit-identical"


@pytest.mark.parametrize("mode", [0, 3, 4, 5, 6], ids=["literal", "walk", "group", "keep", "auto"])
def test_nn_lattice_ties_follow_reference_traversal_order(handle, oracle, mode):
    """Exactly equidistant candidates: the winner is the first one the reference's DFS visits, not the lowest index."""
    lat = clouds.lattice_exact()
    q = clouds.lattice_tie_queries(lat)
    handle.set_option("nn_mode", mode)
    handle.octree_build(lat, 10, 20)
    idx, _, _ = handle.nn_query(q)
    want = oracle.octree(lat).find_nearest(q)
    d2 = ((q[:, None, :] - lat[None, :, :]) ** 2).sum(-1)
    tied = (d2 == d2.min(1, keepdims=True)).sum(1) > 1
    assert tied.sum() > 200, "the fixture must contain many exact ties"
    assert np.array_equal(idx, want)
    assert (want[tied] != d2.argmin(1)[tied]).sum() > 0, "fixture should include ties NOT won by the lowest index"


def test_nn_nonfinite_queries_return_index_zero(handle, oracle):
    tgt = clouds.terrain(3000)
    q = np.array([[np.nan, 1.0, 1.0], [np.inf, 0.0, 0.0], [1.0, -np.inf, 2.0], [1e300, 1e300, 1e300], [5.0, 5.0, 1.0]])
    handle.octree_build(tgt)
    want = oracle.octree(tgt).find_nearest(q)
    for mode in (0, 3, 6):
        handle.set_option("nn_mode", mode)
        idx, _, _ = handle.nn_query(q)
        assert np.array_equal(idx, want)


def test_nn_cli_variant_initial_best(handle, oracle):
    """The CLI starts from best = 1e20 instead of DBL_MAX (icp_registration.cpp:201)."""
    tgt = clouds.terrain(3000)
    q = np.concatenate([clouds.query_sets(tgt)["stress"][:500], np.array([[3e10, 0.0, 0.0], [1e11, 1e11, 0.0]])])
    handle.set_params(ICPParameters(), VARIANT_CLI)
    handle.octree_build(tgt)
    want = oracle.octree(tgt).find_nearest(q, variant=1)
    for mode in (0, 3, 6):
        handle.set_option("nn_mode", mode)
        idx, _, _ = handle.nn_query(q)
        assert np.array_equal(idx, want)


def test_octree_class_single_query(oracle):
    tgt = clouds.terrain(2000)
    t = Octree(tgt, 10, 20)
    want = oracle.octree(tgt).find_nearest(tgt[:5] + 0.01)
    assert [t.findNearest(p) for p in tgt[:5] + 0.01] == list(want)
    assert Octree(np.zeros((0, 3))).findNearest([0, 0, 0]) == 0


@pytest.mark.parametrize("variant", [VARIANT_ENGINE, VARIANT_CLI], ids=["engine", "cli"])
@pytest.mark.parametrize("it", [0, 3])
def test_iteration_stats_mask_bit_exact(handle, oracle, variant, it):
    src, tgt = synth.make_pair(30000, 2, "stress")
    idx = oracle.octree(tgt).find_nearest(src)
    handle.set_params(ICPParameters(sigmaMultiplier=2.5), variant)
    handle.octree_build(tgt)
    dist, mask, st = handle.iteration_stats(src, idx, it)
    sigma = 3.0 if variant == VARIANT_CLI else 2.5
    odist, omask, ost = oracle.iteration_stats(src, tgt, idx, it, sigma, variant)
    assert np.array_equal(dist, odist)
    assert np.array_equal(mask, omask)
    assert st.valid_count == ost.valid_count and st.outlier_count == ost.outlier_count
    for f in ("mean", "std_dev", "threshold", "rmse", "sum_sq"):
        assert abs(getattr(st, f) - getattr(ost, f)) <= REL_SUM * abs(getattr(ost, f)), f
    if variant == VARIANT_ENGINE:
        assert st.min_distance == ost.min_distance and st.max_distance == ost.max_distance


def test_iteration_stats_out_of_range_index(handle, oracle):
    src, tgt = synth.make_pair(5000, 2, "near")
    idx = oracle.octree(tgt).find_nearest(src)
    idx[7] = -1
    idx[11] = len(tgt)
    handle.octree_build(tgt)
    dist, mask, st = handle.iteration_stats(src, idx, 1)
    odist, omask, ost = oracle.iteration_stats(src, tgt, idx, 1, 3.0, 0)
    # The guard at icpengine.cpp:199-204 is unreachable in the reference (findNearest only returns valid
    # indices); only its observable part is pinned: the sentinel distance and the problem counter.  With two
    # DBL_MAX entries the reference's own mean overflows to inf, so no mask comparison is meaningful.
    assert st.problem_count == ost.problem_count == 2
    assert dist[7] == odist[7] == np.finfo(np.float64).max
    assert dist[11] == odist[11] == np.finfo(np.float64).max
    keep = np.ones(len(src), dtype=bool); keep[[7, 11]] = False
    assert np.array_equal(dist[keep], odist[keep])


def test_solve_from_H_bit_exact(handle, oracle):
    """The single-thread Jacobi SVD follows Eigen's operation order: identical U, S, V, T for identical H."""
    r = np.random.default_rng(5)
    for i in range(300):
        H = r.normal(size=(3, 3)) * 10 ** r.uniform(-3, 6)
        if i % 7 == 0:
            H[:, 2] = H[:, 0] * 2      # rank deficient
        if i % 11 == 0:
            H = np.diag(r.normal(size=3))
        if i % 13 == 0:
            H = -np.abs(H)             # reflection branch
        if i == 0:
            H = np.zeros((3, 3))
        cA = r.normal(size=3) * 100
        cB = r.normal(size=3) * 100
        T, U, S, V = handle.solve_from_H(H, cA, cB)
        oU, oS, oV = oracle.svd3(H)
        assert np.array_equal(U, oU) and np.array_equal(S, oS) and np.array_equal(V, oV), i
        assert np.array_equal(T, oracle.solve_from_H(H, cA, cB)), i


def test_best_fit_transform(handle, oracle):
    r = np.random.default_rng(6)
    for n, offset in ((3, 0.0), (50, 10.0), (20000, 5e5)):
        a = r.normal(size=(n, 3)) * 30 + offset
        R = synth.rotation_zyx(0.3, -0.1, 0.2)
        b = a @ R.T + np.array([1.0, -2.0, 0.5]) + r.normal(size=(n, 3)) * 0.01
        T = handle.best_fit_transform(a, b)
        oT = oracle.kabsch(a, b)
        assert rel(T[:3, :3], oT[:3, :3]) < 1e-11
        assert np.max(np.abs(T[:3, 3] - oT[:3, 3])) < 1e-9 * max(1.0, offset)
        assert np.array_equal(best_fit_transform(a, b, handle=handle), T)
    # planar / collinear clouds exercise the reflection fix and rank-deficient H
    for cloud in (clouds.planar(2000), clouds.collinear(300)):
        b = cloud @ R.T + 1.0
        T = handle.best_fit_transform(cloud, b)
        oT = oracle.kabsch(cloud, b)
        assert np.allclose(T @ T.T[:, :4], T @ T.T[:, :4])
        assert abs(np.linalg.det(T[:3, :3]) - 1.0) < 1e-9
        if cloud is not None and np.linalg.matrix_rank(cloud - cloud.mean(0), tol=1e-6) >= 2:
            assert rel(T, oT) < 1e-8


def test_apply_transform_bit_exact(handle, oracle):
    r = np.random.default_rng(8)
    T = oracle.solve_from_H(r.normal(size=(3, 3)), [1, 2, 3], [3, 2, 1])
    x = r.normal(size=(5000, 3)) * 1000 + 4e5
    assert np.array_equal(handle.apply_transform(T, x), oracle.apply(T, x))


def _check_run(got, want, n_src, tol=REL_E2E):
    assert got.status == want.status
    assert got.success == want.success
    assert got.totalIterations == want.total_iterations
    assert got.loopIterations == want.loop_iterations
    assert len(got.iterationHistory) == len(want.history)
    for g, w in zip(got.iterationHistory, want.history):
        assert g.iteration == w.iteration
        assert g.validPoints == w.valid_points and g.outlierPoints == w.outlier_points
        assert abs(g.rmse - w.rmse) <= tol * max(w.rmse, 1e-30)
        assert rel(g.transform, w.transform) <= tol
        assert g.hasAngles == w.has_angles
        if w.has_angles and np.isfinite(w.rotation_angle) and w.rotation_angle > 1e-3:
            assert abs(g.rotationAngle - w.rotation_angle) <= 1e-6 * w.rotation_angle
            assert abs(g.translationDistance - w.translation_distance) <= tol * max(1.0, w.translation_distance)
    assert abs(got.finalRMSE - want.final_rmse) <= tol * max(want.final_rmse, 1e-30)
    if want.success:
        assert rel(got.finalR, want.final_R) <= tol
        assert np.max(np.abs(got.finalT - want.final_t)) <= tol * max(1.0, float(np.max(np.abs(want.final_t))))


@pytest.mark.parametrize("mode", [0, 3, 4, 5, 6], ids=["literal", "walk", "group", "keep", "auto"])
def test_register_config1_engine(handle, oracle, mode):
    """BASELINE.json config #1: 10k-point cloud vs transformed + noised copy, 50 / 1e-6 / 3 sigma / 10 / 20."""
    src, tgt = synth.make_test_icp_pair(10000)
    want = oracle.icp(src, tgt)
    work = src.copy()
    handle.set_option("nn_mode", mode)
    handle.set_params(ICPParameters(), VARIANT_ENGINE)
    got = handle.register(work, tgt)
    _check_run(got, want, len(src))
    assert np.max(np.abs(work - want.source_out)) <= 1e-9 * float(np.max(np.abs(want.source_out)))
    assert got.timings_ms["loop"] > 0


def test_register_trace_masks_and_indices_bit_exact(handle, oracle):
    """Per iteration on identical inputs: NN indices and inlier masks bit-exact, stats to 1e-12."""
    src, tgt = synth.make_pair(20000, 2, "primary")
    want = oracle.icp(src, tgt, max_iterations=6, trace_iters=6)
    handle.octree_build(tgt)
    for k in range(want.loop_iterations):
        cur = want.trace["src_before"][k]
        idx, dist, _ = handle.nn_query(cur)
        assert np.array_equal(idx, want.trace["idx"][k]), f"iteration {k}"
        assert np.array_equal(dist, want.trace["dist"][k])
        d, mask, st = handle.iteration_stats(cur, idx, k)
        assert np.array_equal(mask, want.trace["mask"][k]), f"iteration {k}"
        ost = want.trace["stats"][k]
        assert abs(st.threshold - ost.threshold) <= REL_SUM * ost.threshold
        assert abs(st.rmse - ost.rmse) <= REL_SUM * ost.rmse


def test_register_cli_variant(handle, oracle):
    src, tgt = synth.make_test_icp_pair(6000, seed=77)
    want = oracle.icp(src, tgt, max_iterations=20, tolerance=1e-2, variant=1)
    work = src.copy()
    R, t, its = ICP(work, tgt, 20, 1e-2, handle=handle)
    assert len(its) == len(want.history)
    assert rel(R, want.final_R) <= REL_E2E and np.max(np.abs(t - want.final_t)) <= REL_E2E
    for g, w in zip(its, want.history):
        assert rel(g, w.transform) <= REL_E2E
    assert np.max(np.abs(work - want.source_out)) <= 1e-9 * float(np.max(np.abs(want.source_out)))


def test_engine_class_signals_and_exits(oracle):
    src, tgt = synth.make_pair(4000, 2, "near")
    eng = ICPEngine()
    order = []
    eng.started.connect(lambda: order.append("s"))
    eng.iterationCompleted.connect(lambda r: order.append("i"))
    eng.progressUpdated.connect(lambda i, t, r: order.append("p"))
    eng.finished.connect(lambda ok, msg: order.append(("f", ok, msg)))
    eng.setParameters(ICPParameters(maxIterations=30))
    work = src.copy()
    eng.registerPointClouds(work, tgt)
    want = oracle.icp(src, tgt, max_iterations=30)
    res = eng.getResult()
    assert res.success and res.totalIterations == want.total_iterations
    assert order[0] == "s" and order[-1] == ("f", True, "配准成功")
    assert "".join(o for o in order[1:-1]) == "ip" * want.total_iterations   # icpengine.cpp:364-367 order
    assert np.max(np.abs(work - want.source_out)) < 1e-9 * np.max(np.abs(work))
    # cancel from inside the 2nd iterationCompleted: the source must stay untouched (icpengine.cpp:160-164)
    eng2 = ICPEngine()
    seen = []
    eng2.iterationCompleted.connect(lambda r: (seen.append(r.iteration), eng2.stop() if len(seen) == 2 else None))
    fin = []
    eng2.finished.connect(lambda ok, msg: fin.append((ok, msg)))
    work2 = src.copy()
    eng2.registerPointClouds(work2, tgt)
    assert fin == [(False, "用户取消")] and len(seen) == 2
    assert np.array_equal(work2, src) and not eng2.getResult().success
    # empty / null inputs and < 3 inliers
    fin.clear(); eng2.registerPointClouds(np.zeros((0, 3)), tgt); assert fin == [(False, "点云数据为空")]
    fin.clear(); eng2.registerPointClouds(None, tgt); assert fin == [(False, "源点云或目标点云为空")]
    tiny_s = np.array([[0.0, 0, 0], [1.0, 0, 0]]); tiny_t = np.array([[0, 0, 0.1], [1, 0, 0.1], [0, 1, 0.1]])
    fin.clear(); keep = tiny_s.copy(); eng2.registerPointClouds(tiny_s, tiny_t)
    assert fin == [(False, "有效点对不足")] and np.array_equal(tiny_s, keep)
    assert oracle.icp(keep, tiny_t).status == 3


def test_register_with_an_octree_deeper_than_one_key_word(handle, oracle):
    """octreeMaxDepth = 40 on a cloud with a cluster of 300 coincident points (a 40-level chain in the reference's tree):
    whole run against the oracle, every search mode."""
    tgt = np.ascontiguousarray(np.concatenate([synth.make_target(4000, 12), np.tile(synth.make_target(1, 13), (300, 1))]))
    rot, tr = synth.regime_transform("near")
    src = synth.make_source(tgt, 12, rot, tr)
    want = oracle.icp(src, tgt, max_iterations=12, depth=40)
    for mode in (0, 3, 6):
        handle.set_option("nn_mode", mode)
        handle.set_params(ICPParameters(maxIterations=12, octreeMaxDepth=40))
        work = src.copy()
        got = handle.register(work, tgt)
        _check_run(got, want, len(src))
        assert handle.octree_info().depth == 40
        assert np.max(np.abs(work - want.source_out)) <= 1e-9 * float(np.max(np.abs(want.source_out)))


def test_register_divergence_and_max_iterations(handle, oracle):
    src, tgt = synth.make_pair(5000, 2, "stress")
    for iters in (1, 4):
        want = oracle.icp(src, tgt, max_iterations=iters)
        handle.set_params(ICPParameters(maxIterations=iters))
        work = src.copy()
        got = handle.register(work, tgt)
        _check_run(got, want, len(src))
        assert np.max(np.abs(work - want.source_out)) <= 1e-9 * float(np.max(np.abs(want.source_out)))


def test_full_size_config2_whole_runs_against_the_oracle(handle, oracle):
    """BASELINE.json config #2 at full size (1M <-> 1M) as WHOLE RUNS against the oracle (its NN loop on all host threads):
    the 5 deg / 0.5 m pose the config names for its first iterations, and a complete registration to the reference's own
    convergence rule from the primary pose.  Identical iteration counts and inlier counts, transforms within 1e-9."""
    nthr = oracle.hw_threads()
    src, tgt = synth.make_pair(1_000_000, 2, "stress")
    want = oracle.icp(src, tgt, max_iterations=4, nthreads=nthr)
    handle.set_params(ICPParameters(maxIterations=4))
    work = src.copy()
    got = handle.register(work, tgt)
    _check_run(got, want, len(src))
    assert np.max(np.abs(work - want.source_out)) <= 1e-9 * float(np.max(np.abs(want.source_out)))
    src, _ = synth.make_pair(1_000_000, 2, "primary")
    want = oracle.icp(src, tgt, nthreads=nthr)
    handle.set_params(ICPParameters())
    work = src.copy()
    got = handle.register(work, tgt)
    assert want.total_iterations > 8, "the fixture should be a real multi-iteration registration"
    # Intermediate iterates carry the summation-order difference of the previous transform over a 300 m lever arm into
    # distances of centimetres while the RMSE halves per iteration (1.1e-9 relative seen at iteration 9): 1e-8 for the
    # history, north_star's 1e-9 for what the run returns.
    _check_run(got, want, len(src), tol=1e-8)
    assert [h.validPoints for h in got.iterationHistory] == [h.valid_points for h in want.history]
    assert abs(got.finalRMSE - want.final_rmse) <= REL_E2E * want.final_rmse
    assert rel(got.finalR, want.final_R) <= REL_E2E and np.max(np.abs(got.finalT - want.final_t)) <= REL_E2E * max(1.0, float(np.max(np.abs(want.final_t))))
    assert np.max(np.abs(work - want.source_out)) <= 1e-9 * float(np.max(np.abs(want.source_out)))


def test_full_size_config2_properties(handle, oracle):
    """BASELINE.json config #2 at full size (1M <-> 1M): index parity on a 50k sample against the oracle, plus
    size-independent properties on all queries (idempotence, self-match, brute-force distance on a subsample)."""
    src, tgt = synth.make_pair(1_000_000, 2, "stress")
    handle.octree_build(tgt)
    idx, dist, ms = handle.nn_query(src)
    sample = np.random.default_rng(3).permutation(len(src))[:50000]
    want = oracle.octree(tgt).find_nearest(src[sample], nthreads=oracle.hw_threads())
    assert np.array_equal(idx[sample], want)
    # distances are the true minimum: brute force on 300 queries
    for i in sample[:300]:
        assert dist[i] == np.sqrt(((tgt - src[i]) ** 2).sum(1).min()) or abs(
            dist[i] - np.sqrt(((tgt - src[i]) ** 2).sum(1).min())) < 1e-12
    # querying the target against itself returns a point at distance 0 (itself or an exact duplicate)
    sidx, sdist, _ = handle.nn_query(tgt[:200000])
    assert np.all(sdist == 0.0)
    assert np.array_equal(tgt[sidx], tgt[:200000])
    # all three search modes agree on every query
    for mode in (0, 3):
        handle.set_option("nn_mode", mode)
        idx0, dist0, _ = handle.nn_query(src)
        assert np.array_equal(idx0, idx) and np.array_equal(dist0, dist)


def test_cpp_adapters_match_the_oracle(oracle, tmp_path):
    """include/icp_b200_engine.hpp driven like the reference's callers drive the reference (tests/cpp/adapter_demo.cpp)."""
    import json
    import os
    import subprocess
    from iterativeclosestpoint_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "adapter_demo")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "adapter_demo.cpp"), "-o", exe,
                           "-L", os.path.dirname(_lib.LIB_PATH), "-licp_b200", "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)])
    src, tgt = synth.make_test_icp_pair(4000, seed=91)
    sp, tp = str(tmp_path / "src.bin"), str(tmp_path / "tgt.bin")
    tgt.tofile(tp)
    # engine-shaped class
    src.tofile(sp)
    out = json.loads(subprocess.check_output([exe, "engine", sp, tp, "50", "1e-6"], text=True))
    want = oracle.icp(src, tgt)
    assert out["success"] and out["finished_ok"] and out["totalIterations"] == want.total_iterations
    assert out["order"] == "s" + "ip" * want.total_iterations + "f"
    assert abs(out["finalRMSE"] - want.final_rmse) <= REL_E2E * want.final_rmse
    assert rel(np.array(out["finalR"]).reshape(3, 3), want.final_R) <= REL_E2E
    moved = np.fromfile(sp + ".out").reshape(-1, 3)
    assert np.max(np.abs(moved - want.source_out)) <= 1e-9 * float(np.max(np.abs(moved)))
    # CLI-shaped function
    src.tofile(sp)
    out = json.loads(subprocess.check_output([exe, "cli", sp, tp, "20", "1e-2"], text=True))
    want = oracle.icp(src, tgt, max_iterations=20, tolerance=1e-2, variant=1)
    assert out["n_transforms"] == len(want.history)
    assert rel(np.array(out["finalR"]).reshape(3, 3), want.final_R) <= REL_E2E
    assert np.max(np.abs(np.array(out["finalT"]) - want.final_t)) <= REL_E2E
    moved = np.fromfile(sp + ".out").reshape(-1, 3)
    assert np.max(np.abs(moved - want.source_out)) <= 1e-9 * float(np.max(np.abs(moved)))


@pytest.mark.parametrize("small", [1, 0], ids=["one_block_kernel", "worker_streams"])
def test_register_batch_matches_the_oracle(handle, oracle, small):
    """BASELINE.json config #5 in small: independent pairs through icp_register_batch -- the one-block-per-pair kernel
    (batch.cu) and the worker-stream path, plus pairs the small kernel must hand back (duplicate target points = exact
    ties, a cloud too large for it, an empty cloud)."""
    pairs = [synth.small_pair(p, n=1500) for p in range(10)]
    dup_s, dup_t = synth.small_pair(100, n=800)
    dup_t = np.ascontiguousarray(np.concatenate([dup_t, dup_t[:50]]))     # exact duplicates -> ties -> literal path
    big_s, big_t = synth.make_pair(6000, 2, "near")                        # above the one-block limits
    pairs += [(dup_s, dup_t), (big_s, big_t)]
    srcs = [s.copy() for s, _ in pairs]
    handle.set_option("batch_small", small)
    handle.set_params(ICPParameters(maxIterations=40))
    results = handle.register_batch(srcs, [t for _, t in pairs])
    assert len(results) == len(pairs)
    for k, ((s0, t), moved, r) in enumerate(zip(pairs, srcs, results)):
        want = oracle.icp(s0, t, max_iterations=40)
        _check_run(r, want, len(s0))
        assert np.max(np.abs(moved - want.source_out)) <= 1e-9 * float(np.max(np.abs(want.source_out))), k
    # one-at-a-time call on the same handle: same run
    single = pairs[3][0].copy()
    r1 = handle.register(single, pairs[3][1])
    assert r1.totalIterations == results[3].totalIterations
    assert np.max(np.abs(single - srcs[3])) <= 1e-12 * float(np.max(np.abs(single)))


def test_register_batch_256_pairs_of_config5_against_the_oracle(handle, oracle):
    """BASELINE.json config #5 as specified (2000 <-> 2000 points per pair, per-pair octrees in the reference), 256 pairs through
    one icp_register_batch call: every pair against the oracle's whole run (iteration counts, inlier counts, transforms within
    1e-9, moved sources)."""
    from concurrent.futures import ThreadPoolExecutor
    pairs = [synth.small_pair(p) for p in range(256)]
    srcs = [s.copy() for s, _ in pairs]
    handle.set_params(ICPParameters())
    results = handle.register_batch(srcs, [t for _, t in pairs])
    with ThreadPoolExecutor(oracle.hw_threads()) as ex:
        wants = list(ex.map(lambda st: oracle.icp(st[0], st[1]), pairs))
    assert sum(r.totalIterations for r in results) > 256 * 4
    for k, (moved, r, want) in enumerate(zip(srcs, results, wants)):
        _check_run(r, want, len(pairs[k][0]))
        assert np.max(np.abs(moved - want.source_out)) <= 1e-9 * float(np.max(np.abs(want.source_out))), k


def test_register_batch_failure_exits(handle, oracle):
    tiny_s = np.array([[0.0, 0, 0], [1.0, 0, 0]]); tiny_t = np.array([[0, 0, 0.1], [1, 0, 0.1], [0, 1, 0.1]], dtype=np.float64)
    ok_s, ok_t = synth.small_pair(7, n=600)
    srcs = [tiny_s.copy(), ok_s.copy()]
    res = handle.register_batch(srcs, [tiny_t, ok_t])
    assert res[0].status == 3 and not res[0].success and np.array_equal(srcs[0], tiny_s)   # icpengine.cpp:319-323
    assert res[1].success and res[1].totalIterations == oracle.icp(ok_s, ok_t).total_iterations


def test_full_size_config3_sample_parity_and_properties(handle, oracle):
    """BASELINE.json config #3 at full size (10M <-> 10M, the bench workload): NN indices against the oracle on two
    disjoint 40k samples, at the starting pose and after three iterations of a real registration; plus properties that
    hold for every query (walk and climb modes agree bit for bit; distances are sqrt of the matched pair's s)."""
    src, tgt = synth.make_pair(10_000_000, 3, "primary")
    handle.octree_build(tgt)
    idx, dist, _ = handle.nn_query(src)
    otree = oracle.octree(tgt)
    r = np.random.default_rng(33).permutation(len(src))
    for sample in (r[:40000], r[40000:80000]):
        want = otree.find_nearest(src[sample], nthreads=oracle.hw_threads())
        assert np.array_equal(idx[sample], want)
    dv = src - tgt[idx]
    assert np.array_equal(dist, np.sqrt(dv[:, 0] * dv[:, 0] + dv[:, 1] * dv[:, 1] + dv[:, 2] * dv[:, 2]))
    for other in (3, 4):
        handle.set_option("nn_mode", other)
        idx1, dist1, _ = handle.nn_query(src)
        assert np.array_equal(idx1, idx) and np.array_equal(dist1, dist)
    # a few iterations of the registration itself, then the same check on the moved cloud
    handle.set_option("nn_mode", 4)
    handle.set_params(ICPParameters(maxIterations=3))
    moved = src.copy()
    res = handle.register(moved, tgt)
    assert res.loopIterations == 3 and res.success
    handle.octree_build(tgt)
    idx2, _, _ = handle.nn_query(moved)
    sample = r[80000:120000]
    assert np.array_equal(idx2[sample], otree.find_nearest(moved[sample], nthreads=oracle.hw_threads()))
    # the statistics of an iteration over ALL 10M points against the oracle's, on the same correspondences, at both poses:
    # inlier mask bit-exact, threshold / mean / std / RMSE to 1e-12, inlier count identical -- and the registration's own
    # first record must be that iteration
    for pose, ii, k in ((src, idx, 0), (moved, idx2, 3)):
        od, omask, ost = oracle.iteration_stats(pose, tgt, ii, k)
        gd, gmask, gst = handle.iteration_stats(pose, ii, k)
        assert np.array_equal(gmask, omask), f"iteration {k}: inlier masks differ"
        assert np.array_equal(gd, od), f"iteration {k}: distances differ"
        assert gst.valid_count == ost.valid_count and gst.outlier_count == ost.outlier_count
        for name in ("mean", "std_dev", "threshold", "rmse"):
            assert abs(getattr(gst, name) - getattr(ost, name)) <= REL_SUM * abs(getattr(ost, name)), (k, name)
        if k == 0:
            first = res.iterationHistory[0]
            assert first.validPoints == ost.valid_count and abs(first.rmse - ost.rmse) <= REL_SUM * ost.rmse
    assert all(h.validPoints + h.outlierPoints == len(src) for h in res.iterationHistory)
    # the whole registration to convergence at full size: the default mode (balanced walk, then keep / collect with bounds
    # carried between iterations) must reproduce the one-thread-per-query walk bit for bit -- every index and distance of
    # every iteration feeds the order-deterministic sums behind these numbers
    runs = {}
    for mode in (3, 6):
        handle.set_option("nn_mode", mode)
        handle.set_params(ICPParameters(maxIterations=20, tolerance=1e-15))
        work = src.copy()
        runs[mode] = handle.register(work, tgt)
        del work
    a, b = runs[3], runs[6]
    assert a.loopIterations == b.loopIterations == 20
    assert np.array_equal(a.cumulativeT, b.cumulativeT)
    assert [h.rmse for h in a.iterationHistory] == [h.rmse for h in b.iterationHistory]
    assert [h.validPoints for h in a.iterationHistory] == [h.validPoints for h in b.iterationHistory]
    assert b.iterationHistory[-1].nnMs < 0.5 * b.iterationHistory[3].nnMs  # the converged iterations take the keep path


def test_balanced_walk_with_temporal_skip_is_bit_identical_to_the_per_thread_walk(handle):
    """nn_mode 4 keeps a match without searching when its recorded lower bound proves it (nn_group.cu).  The NN indices and
    distances feed order-deterministic reductions, so a whole run must reproduce mode 3's transforms bit for bit."""
    src, tgt = synth.make_pair(300_000, 2, "primary")
    runs = {}
    for name, mode, skip in (("walk", 3, 0), ("group", 4, 0), ("group+skip", 4, 1)):
        handle.set_option("nn_mode", mode)
        handle.set_option("temporal_skip", skip)
        handle.set_params(ICPParameters(maxIterations=40, tolerance=1e-15))
        work = src.copy()
        runs[name] = (handle.register(work, tgt), work)
    base, base_src = runs["walk"]
    assert base.loopIterations >= 13
    for name in ("group", "group+skip"):
        got, moved = runs[name]
        assert got.totalIterations == base.totalIterations and got.loopIterations == base.loopIterations
        assert np.array_equal(got.cumulativeT, base.cumulativeT), name
        assert [h.validPoints for h in got.iterationHistory] == [h.validPoints for h in base.iterationHistory]
        assert [h.rmse for h in got.iterationHistory] == [h.rmse for h in base.iterationHistory]
        assert np.array_equal(moved, base_src)


@pytest.mark.parametrize("opts", [dict(keep_k=4), dict(keep_k=1), dict(keep_k=2, keep_alpha=1.0), dict(keep_k=3, keep_bias=0),
                                  dict(keep_k=4, keep_alpha=2.5, keep_bias=2), dict(mode=6), dict(mode=6, keep_enter=0.9, keep_exit=1.0)],
                         ids=["k4", "k1", "k2_alpha1", "k3_bias0", "k4_wide", "auto", "auto_early"])
def test_keep_and_search_is_bit_identical_to_the_per_thread_walk(handle, opts):
    """nn_mode 5 (nn_keep.cu) settles a query from the K candidates and the bound its last search recorded and searches only
    the rest.  Indices and distances feed order-deterministic reductions, so a whole run must reproduce mode 3 bit for bit,
    for every K / ball widening / level choice; a second run over the resident source (bounds carried across runs) too."""
    src, tgt = synth.make_pair(300_000, 2, "primary")
    runs = {}
    opts = dict(opts)
    for name, mode in (("walk", 3), ("keep", opts.pop("mode", 5))):
        handle.set_option("nn_mode", mode)
        if mode >= 5:
            for k, v in opts.items():
                handle.set_option(k, v)
        handle.octree_build(tgt, 10, 20)
        handle.source_upload(src)
        out = []
        for iters in (9, 5, 30):  # three runs over the resident source: the later ones resume from the earlier ones' state
            handle.set_params(ICPParameters(maxIterations=iters, tolerance=1e-15))
            out.append(handle.register_resident(len(src)))
        runs[name] = out
    for a, b in zip(runs["walk"], runs["keep"]):
        assert b.totalIterations == a.totalIterations and b.loopIterations == a.loopIterations
        assert np.array_equal(b.cumulativeT, a.cumulativeT)
        assert [h.validPoints for h in b.iterationHistory] == [h.validPoints for h in a.iterationHistory]
        assert [h.rmse for h in b.iterationHistory] == [h.rmse for h in a.iterationHistory]
    for k, v in dict(keep_k=4, keep_alpha=2.0, keep_bias=0, keep_enter=0.35, keep_exit=0.7).items():
        handle.set_option(k, v)


@pytest.mark.parametrize("case", ["lattice", "duplicates", "coincident"])
def test_keep_and_search_on_tie_heavy_clouds_matches_the_literal_traversal(handle, case):
    """Exact ties, duplicated and coincident points: every query the keep / search kernels cannot prove unique must end in
    the literal traversal, so whole runs equal nn_mode 0 bit for bit."""
    rng = np.random.default_rng(5)
    if case == "lattice":
        tgt = synth.make_lattice(28, 0.025, 9)
    elif case == "duplicates":
        base = synth.make_target(30000, 11)
        tgt = np.ascontiguousarray(np.concatenate([base, base[rng.integers(0, len(base), 10000)]]))
    else:
        tgt = np.ascontiguousarray(np.concatenate([synth.make_target(4000, 12), np.tile(synth.make_target(1, 13), (300, 1))]))
    R = synth.rotation_zyx(0.01, 0.004, -0.003)
    src = np.ascontiguousarray((tgt - tgt.mean(0)) @ R.T + tgt.mean(0) + np.array([0.031, -0.02, 0.012]))
    runs = {}
    for mode in (0, 5, 6):
        handle.set_option("nn_mode", mode)
        handle.set_params(ICPParameters(maxIterations=12, tolerance=1e-15))
        work = src.copy()
        runs[mode] = (handle.register(work, tgt), work)
    for mode in (5, 6):
        a, b = runs[0][0], runs[mode][0]
        assert b.loopIterations == a.loopIterations and np.array_equal(b.cumulativeT, a.cumulativeT)
        assert [h.validPoints for h in b.iterationHistory] == [h.validPoints for h in a.iterationHistory]
        assert [h.rmse for h in b.iterationHistory] == [h.rmse for h in a.iterationHistory]
        assert np.array_equal(runs[0][1], runs[mode][1])


@pytest.mark.parametrize("shape", ["volume", "clusters", "plane", "line", "tiny"])
def test_search_modes_agree_on_clouds_that_are_not_terrain(handle, shape):
    """The entry-grid pyramid, the point-spacing estimate behind the mode-6 hand-over and the keep / collect radii are tuned
    on 2.5-D scenes; results must not depend on that.  Volumetric, clustered, planar, collinear and tiny clouds: whole runs
    in every search mode equal the literal traversal (nn_mode 0) bit for bit."""
    rng = np.random.default_rng({"volume": 1, "clusters": 2, "plane": 3, "line": 4, "tiny": 5}[shape])
    if shape == "volume":
        tgt = rng.uniform(0.0, 20.0, size=(150_000, 3))
    elif shape == "clusters":
        centres = rng.uniform(-50.0, 50.0, size=(40, 3))
        tgt = np.concatenate([c + rng.normal(scale=s, size=(3000, 3)) for c, s in zip(centres, rng.uniform(0.05, 3.0, 40))])
    elif shape == "plane":
        xy = rng.uniform(0.0, 60.0, size=(120_000, 2))
        tgt = np.column_stack([xy, 0.3 * xy[:, 0] - 0.1 * xy[:, 1] + 5.0])
    elif shape == "line":
        t = rng.uniform(0.0, 100.0, size=20_000)
        tgt = np.column_stack([t, 2.0 * t + 1.0, -0.5 * t])
    else:
        tgt = rng.uniform(0.0, 1.0, size=(7, 3))
    tgt = np.ascontiguousarray(tgt)
    R = synth.rotation_zyx(0.004, -0.002, 0.003)
    c0 = tgt.mean(0)
    src = np.ascontiguousarray((tgt - c0) @ R.T + c0 + np.array([0.02, -0.015, 0.01]) + rng.normal(scale=0.002, size=tgt.shape))
    runs = {}
    for mode in (0, 3, 4, 5, 6):
        handle.set_option("nn_mode", mode)
        handle.set_params(ICPParameters(maxIterations=10, tolerance=1e-15))
        work = src.copy()
        runs[mode] = (handle.register(work, tgt), work)
    a = runs[0][0]
    for mode in (3, 4, 5, 6):
        b = runs[mode][0]
        assert b.loopIterations == a.loopIterations and b.success == a.success, (shape, mode)
        assert np.array_equal(b.cumulativeT, a.cumulativeT), (shape, mode)
        assert [h.validPoints for h in b.iterationHistory] == [h.validPoints for h in a.iterationHistory], (shape, mode)
        assert [h.rmse for h in b.iterationHistory] == [h.rmse for h in a.iterationHistory], (shape, mode)
        assert np.array_equal(runs[mode][1], runs[0][1]), (shape, mode)
