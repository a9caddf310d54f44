"""Host-side mirror of the reference's interface for the ICP hot path, over the C ABI (include/icp_b200.h).

Names, argument meaning and error behaviour follow the reference so parity tests read like tests of the
reference itself:

    ICPParameters / IterationResult / ICPResult     PointCloudRegistration/core/icpengine.h:13-44
    ICPEngine.setParameters / registerPointClouds /
              stop / getResult + the five signals    core/icpengine.h:60-75
    Octree(pts, max_pts, max_d).findNearest          core/octree.h:27-43
    ICP(source, target, max_iterations, tolerance)   icp_registration.cpp:443-446
    best_fit_transform(A, B)                         icp_registration.cpp:389

Clouds are numpy (n,3) float64 arrays == std::vector<Point3D>; `registerPointClouds` and `ICP` update the
source array IN PLACE exactly where the reference writes its source back.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import IcpError, VARIANT_CLI, VARIANT_ENGINE


def _c3(a, writable=False):
    a = np.asarray(a)
    if a.dtype != np.float64 or not a.flags.c_contiguous or a.ndim != 2 or a.shape[1] != 3:
        if writable:
            raise ValueError("cloud must be a C-contiguous float64 array of shape (n, 3)")
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 3)
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@dataclass
class ICPParameters:  # core/icpengine.h:13-19
    maxIterations: int = 50
    tolerance: float = 1e-6
    sigmaMultiplier: float = 3.0
    octreeMaxPoints: int = 10
    octreeMaxDepth: int = 20


@dataclass
class IterationResult:  # core/icpengine.h:24-32
    iteration: int = 0
    rmse: float = 0.0
    validPoints: int = 0
    outlierPoints: int = 0
    transform: np.ndarray = field(default_factory=lambda: np.eye(4))
    rotationAngle: float = float("nan")
    translationDistance: float = float("nan")
    hasAngles: bool = True
    nnMs: float = 0.0
    iterMs: float = 0.0


@dataclass
class ICPResult:  # core/icpengine.h:37-44
    success: bool = False
    totalIterations: int = 0
    finalRMSE: float = 0.0
    finalR: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    finalT: np.ndarray = field(default_factory=lambda: np.zeros(3))
    iterationHistory: list = field(default_factory=list)
    # extras of this implementation
    status: int = 0
    loopIterations: int = 0
    cumulativeT: np.ndarray = field(default_factory=lambda: np.eye(4))
    lastT: np.ndarray = field(default_factory=lambda: np.eye(4))
    timings_ms: dict = field(default_factory=dict)


class Signal:
    """Stand-in for a Qt signal: connect(callable); emit(*args) calls the slots in connection order."""

    def __init__(self):
        self._slots = []

    def connect(self, fn):
        self._slots.append(fn)

    def emit(self, *a):
        for fn in self._slots:
            fn(*a)


class Handle:
    """One icp_handle (one CUDA device)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        st = self.lib.icp_create(C.byref(self.h), device)
        if st != _lib.ICP_OK:
            self.h = None
            raise IcpError(st, "icp_create failed (no CUDA device? this library has no CPU fallback)")
        self.device = device
        self._cbs = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.icp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st, ok=(0,)):
        if st not in ok:
            raise IcpError(st, (self.lib.icp_last_error(self.h) or b"").decode("utf-8", "replace"))
        return st

    # -- parameters / options ---------------------------------------------------------------------
    def set_params(self, p: ICPParameters, variant: int = VARIANT_ENGINE):
        cp = _lib.IcpParams(p.maxIterations, p.octreeMaxPoints, p.octreeMaxDepth, variant, p.tolerance, p.sigmaMultiplier)
        self.check(self.lib.icp_set_params(self.h, C.byref(cp)))

    def set_option(self, key: str, value: float):
        self.check(self.lib.icp_set_option(self.h, key.encode(), float(value)))

    def set_callbacks(self, on_iteration=None, on_progress=None, on_log=None):
        def _it(p, _u):
            if on_iteration:
                on_iteration(_iteration_from_c(p.contents))

        def _pr(i, t, r, _u):
            if on_progress:
                on_progress(i, t, r)

        def _lg(m, _u):
            if on_log:
                on_log(m.decode("utf-8", "replace"))

        self._cbs = (_lib.ITERATION_CB(_it), _lib.PROGRESS_CB(_pr), _lib.LOG_CB(_lg))
        self.check(self.lib.icp_set_callbacks(self.h, self._cbs[0], self._cbs[1], self._cbs[2], None))

    def kernel_launches(self) -> int:
        return int(self.lib.icp_kernel_launches(self.h))

    def nn_counters(self, reset=True):
        """(queries answered by the fast path, queries re-run through the literal traversal) since the last reset."""
        a = C.c_int64(); b = C.c_int64()
        self.check(self.lib.icp_nn_counters(self.h, C.byref(a), C.byref(b), 1 if reset else 0))
        return a.value, b.value

    def nn_tile_counters(self, reset=True):
        """(tile lanes handed to the per-thread search, candidates scanned by tiles) since the last reset."""
        a = C.c_int64(); b = C.c_int64()
        self.check(self.lib.icp_nn_tile_counters(self.h, C.byref(a), C.byref(b), 1 if reset else 0))
        return a.value, b.value

    # -- whole path ---------------------------------------------------------------------------------
    def _result(self, cap):
        hist = (_lib.IcpIteration * max(cap, 1))()
        res = _lib.IcpResult()
        res.history = C.cast(hist, C.POINTER(_lib.IcpIteration))
        res.history_cap = cap
        return res, hist

    def register(self, source, target, max_history=None, stop_flag=None) -> ICPResult:
        p = _lib.IcpParams()
        self.lib.icp_get_params(self.h, C.byref(p))
        res, hist = self._result((max_history or p.max_iterations) + 2)
        src = _c3(source, writable=True) if source is not None else None
        tgt = _c3(target) if target is not None else None
        st = self.lib.icp_register(self.h, _ptr(src), 0 if src is None else len(src), _ptr(tgt),
                                   0 if tgt is None else len(tgt), C.byref(res),
                                   C.byref(stop_flag) if stop_flag is not None else None)
        self.check(st, ok=(0, 1, 2, 3))
        return _result_from_c(res, hist)

    def register_sharded(self, source_shard, n_src_global, target, stop_flag=None) -> ICPResult:
        p = _lib.IcpParams()
        self.lib.icp_get_params(self.h, C.byref(p))
        res, hist = self._result(p.max_iterations + 2)
        src = _c3(source_shard, writable=True)
        tgt = _c3(target)
        st = self.lib.icp_register_sharded(self.h, _ptr(src), len(src), int(n_src_global), _ptr(tgt), len(tgt),
                                           C.byref(res), C.byref(stop_flag) if stop_flag is not None else None)
        self.check(st, ok=(0, 1, 2, 3))
        return _result_from_c(res, hist)

    def register_batch(self, sources, targets) -> list:
        """icp_register_batch: independent pairs (BASELINE.json config #5); every source array is updated in place."""
        n = len(sources)
        srcs = [_c3(a, writable=True) for a in sources]
        tgts = [_c3(a) for a in targets]
        p = _lib.IcpParams()
        self.lib.icp_get_params(self.h, C.byref(p))
        cap = p.max_iterations + 2
        res = (_lib.IcpResult * n)()
        hists = []
        for k in range(n):
            hst = (_lib.IcpIteration * cap)()
            hists.append(hst)
            res[k].history = C.cast(hst, C.POINTER(_lib.IcpIteration))
            res[k].history_cap = cap
        sp = (C.c_void_p * n)(*[a.ctypes.data for a in srcs])
        tp = (C.c_void_p * n)(*[a.ctypes.data for a in tgts])
        ns = (C.c_int64 * n)(*[len(a) for a in srcs])
        nt = (C.c_int64 * n)(*[len(a) for a in tgts])
        import time as _time
        t0 = _time.perf_counter()
        st = self.lib.icp_register_batch(self.h, n, sp, ns, tp, nt, res)
        self.last_batch_seconds = _time.perf_counter() - t0   # the C call alone (result unpacking below is Python overhead)
        self.check(st, ok=(0, 1, 2, 3))
        return [_result_from_c(res[k], hists[k]) for k in range(n)]

    def source_upload(self, source):
        src = _c3(source)
        self.check(self.lib.icp_source_upload(self.h, _ptr(src), len(src)))

    def register_resident(self, n_src_global=0, source_out=None) -> ICPResult:
        p = _lib.IcpParams()
        self.lib.icp_get_params(self.h, C.byref(p))
        res, hist = self._result(p.max_iterations + 2)
        out = _c3(source_out, writable=True) if source_out is not None else None
        st = self.lib.icp_register_resident(self.h, int(n_src_global), C.byref(res), _ptr(out), None)
        self.check(st, ok=(0, 1, 2, 3))
        return _result_from_c(res, hist)

    # -- stages ---------------------------------------------------------------------------------------
    def octree_build(self, target, max_points=10, max_depth=20):
        tgt = _c3(target)
        self.check(self.lib.icp_octree_build(self.h, _ptr(tgt), len(tgt), max_points, max_depth))

    def octree_info(self) -> _lib.IcpOctreeInfo:
        info = _lib.IcpOctreeInfo()
        self.check(self.lib.icp_octree_get_info(self.h, C.byref(info)))
        return info

    def octree_dump(self) -> dict:
        nn = C.c_int64()
        ni = C.c_int64()
        self.check(self.lib.icp_octree_dump(self.h, C.byref(nn), C.byref(ni), None, None, None, None, None, None))
        depth = np.empty(nn.value, dtype=np.int32)
        key = np.empty(nn.value, dtype=np.uint64)
        leaf = np.empty(nn.value, dtype=np.uint8)
        count = np.empty(nn.value, dtype=np.int32)
        box = np.empty((nn.value, 6))
        idx = np.empty(ni.value, dtype=np.int32)
        self.check(self.lib.icp_octree_dump(self.h, C.byref(nn), C.byref(ni), _ptr(depth), _ptr(key), _ptr(leaf),
                                            _ptr(count), _ptr(box), _ptr(idx)))
        return {"depth": depth, "key": key, "leaf": leaf, "count": count, "box": box, "idx": idx}

    def nn_query(self, queries, want_dist=True):
        q = _c3(queries)
        idx = np.empty(len(q), dtype=np.int32)
        dist = np.empty(len(q)) if want_dist else None
        ms = C.c_float(0)
        self.check(self.lib.icp_nn_query(self.h, _ptr(q), len(q), _ptr(idx), _ptr(dist), C.byref(ms)))
        return idx, dist, float(ms.value)

    def iteration_stats(self, source, idx, iteration=0):
        src = _c3(source)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        dist = np.empty(len(src))
        mask = np.empty(len(src), dtype=np.uint8)
        st = _lib.IcpStats()
        self.check(self.lib.icp_iteration_stats(self.h, _ptr(src), len(src), _ptr(idx), iteration, _ptr(dist),
                                                _ptr(mask), C.byref(st)))
        return dist, mask, st

    def best_fit_transform(self, a, b):
        a = _c3(a)
        b = _c3(b)
        T = np.empty(16)
        self.check(self.lib.icp_best_fit_transform(self.h, _ptr(a), _ptr(b), len(a), _ptr(T)))
        return T.reshape(4, 4)

    def solve_from_H(self, H, cA, cB):
        H = np.ascontiguousarray(H, dtype=np.float64).reshape(9)
        cA = np.ascontiguousarray(cA, dtype=np.float64)
        cB = np.ascontiguousarray(cB, dtype=np.float64)
        T = np.empty(16); U = np.empty(9); S = np.empty(3); V = np.empty(9)
        self.check(self.lib.icp_solve_from_H(self.h, _ptr(H), _ptr(cA), _ptr(cB), _ptr(T), _ptr(U), _ptr(S), _ptr(V)))
        return T.reshape(4, 4), U.reshape(3, 3), S, V.reshape(3, 3)

    def apply_transform(self, T, xyz):
        out = _c3(xyz).copy()
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        self.check(self.lib.icp_apply_transform(self.h, _ptr(T), _ptr(out), len(out)))
        return out

    # -- multi-GPU ------------------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self.check(self.lib.icp_comm_unique_id(self.h, buf))
        return buf.raw

    def comm_init(self, rank: int, n_ranks: int, unique_id: bytes):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self.check(self.lib.icp_comm_init(self.h, rank, n_ranks, buf))


def _iteration_from_c(it) -> IterationResult:
    r = IterationResult()
    r.iteration = int(it.iteration)
    r.rmse = float(it.rmse)
    r.validPoints = int(it.valid_points)
    r.outlierPoints = int(it.outlier_points)
    r.transform = np.array(list(it.transform), dtype=np.float64).reshape(4, 4)
    r.hasAngles = bool(it.has_angles)
    r.nnMs = float(it.nn_ms)
    r.iterMs = float(it.iter_ms)
    if r.hasAngles:
        r.rotationAngle = float(it.rotation_angle)
        r.translationDistance = float(it.translation_distance)
    return r


def _result_from_c(res, hist) -> ICPResult:
    out = ICPResult()
    out.status = int(res.status)
    out.success = bool(res.success)
    out.totalIterations = int(res.total_iterations)
    out.loopIterations = int(res.loop_iterations)
    out.finalRMSE = float(res.final_rmse)
    out.finalR = np.array(list(res.final_R)).reshape(3, 3)
    out.finalT = np.array(list(res.final_t))
    out.cumulativeT = np.array(list(res.cumulative_T)).reshape(4, 4)
    out.lastT = np.array(list(res.last_T)).reshape(4, 4)
    out.iterationHistory = [_iteration_from_c(hist[k]) for k in range(int(res.history_len))]
    out.timings_ms = {"h2d": res.ms_h2d, "build": res.ms_build, "loop": res.ms_loop, "d2h": res.ms_d2h,
                      "nn_total": res.ms_nn_total, "nn_first": res.ms_nn_first}
    return out


class Octree:
    """Octree(pts, max_pts=10, max_d=20) with findNearest(query) (core/octree.h:27-43)."""

    def __init__(self, pts, max_pts: int = 10, max_d: int = 20, device: int = 0, handle: Handle | None = None):
        self._h = handle or Handle(device)
        self._empty = pts is None or len(pts) == 0
        if not self._empty:
            self._h.octree_build(pts, max_pts, max_d)

    def findNearest(self, query) -> int:
        if self._empty:
            return 0  # octree.cpp:177
        q = np.asarray(query, dtype=np.float64).reshape(1, 3)
        return int(self._h.nn_query(q, want_dist=False)[0][0])

    def find_nearest(self, queries) -> np.ndarray:
        """Batch form of the loop at core/icpengine.cpp:172-184."""
        if self._empty:
            return np.zeros(len(queries), dtype=np.int32)
        return self._h.nn_query(queries, want_dist=False)[0]


class ICPEngine:
    """Same surface as the reference's ICPEngine (core/icpengine.h:51-87); Qt signals become `Signal`s."""

    MSG_NULL = "源点云或目标点云为空"
    MSG_EMPTY = "点云数据为空"
    MSG_CANCEL = "用户取消"
    MSG_FEW = "有效点对不足"
    MSG_OK = "配准成功"

    def __init__(self, device: int = 0):
        self._h = Handle(device)
        self.m_params = ICPParameters()
        self.m_result = ICPResult()
        self._stop = C.c_int(0)
        self.started = Signal()
        self.progressUpdated = Signal()
        self.iterationCompleted = Signal()
        self.finished = Signal()
        self.logMessage = Signal()
        self._h.set_callbacks(self.iterationCompleted.emit, self.progressUpdated.emit, self.logMessage.emit)

    def setParameters(self, params: ICPParameters):
        self.m_params = params

    def getParameters(self) -> ICPParameters:
        return self.m_params

    def stop(self):  # icpengine.cpp:62-66
        self._stop.value = 1
        self.logMessage.emit("用户请求停止配准...")  # icpengine.cpp:65

    def getResult(self) -> ICPResult:
        return self.m_result

    def registerPointClouds(self, source, target):  # icpengine.cpp:24-60
        if source is None or target is None:
            self.finished.emit(False, self.MSG_NULL)
            return
        if len(source) == 0 or len(target) == 0:
            self.finished.emit(False, self.MSG_EMPTY)
            return
        self._stop.value = 0
        self.m_result = ICPResult()
        self.started.emit()
        self._h.set_params(self.m_params, VARIANT_ENGINE)
        self.m_result = self._h.register(source, target, stop_flag=self._stop)
        st = self.m_result.status
        if st == _lib.ICP_OK:
            self.finished.emit(True, self.MSG_OK)
        elif st == _lib.ICP_CANCELLED:
            self.finished.emit(False, self.MSG_CANCEL)
        elif st == _lib.ICP_TOO_FEW_INLIERS:
            self.finished.emit(False, self.MSG_FEW)
        else:
            self.finished.emit(False, self.MSG_EMPTY)


def ICP(source, target, max_iterations: int, tolerance: float, device: int = 0, handle: Handle | None = None):
    """The CLI's ICP() (icp_registration.cpp:443-446): updates `source` in place and returns
    (final_R, final_t, iteration_transforms) -- final_R/final_t are the LAST incremental transform, the list
    holds the cumulative transform of every iteration (:593-595, :616-621)."""
    h = handle or Handle(device)
    h.set_params(ICPParameters(maxIterations=max_iterations, tolerance=tolerance), VARIANT_CLI)
    res = h.register(source, target)
    return res.finalR, res.finalT, [it.transform for it in res.iterationHistory]


def best_fit_transform(A, B, device: int = 0, handle: Handle | None = None) -> np.ndarray:
    """best_fit_transform(A /*N x 3*/, B /*N x 3*/) -> 4x4 (icp_registration.cpp:389-440)."""
    h = handle or Handle(device)
    return h.best_fit_transform(A, B)


def rotation_angle_deg(T) -> float:
    """icpengine.cpp:360-361 (no clamping)."""
    tr = T[0][0] + (T[1][1] + T[2][2])
    v = (tr - 1.0) / 2.0
    return math.degrees(math.acos(v)) if -1.0 <= v <= 1.0 else float("nan")
