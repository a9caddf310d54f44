"""TEST INFRASTRUCTURE ONLY: ctypes bindings for liboracle.so (this repo's CPU restatement, icp_oracle.c)
and, when they were built, oracle/_ref/libref_{engine,cli}.so (the unmodified reference compiled in place).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_ENGINE_SO = os.path.join(HERE, "_ref", "libref_engine.so")
REF_CLI_SO = os.path.join(HERE, "_ref", "libref_cli.so")

VARIANT_ENGINE = 0
VARIANT_CLI = 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _c3(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.ndim == 2 and a.shape[1] == 3
    return a


def build_oracle(force: bool = False) -> str:
    """Compile liboracle.so from icp_oracle.c with the flags of oracle/Makefile (gcc only, seconds)."""
    src = os.path.join(HERE, "icp_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


class OrcStats(C.Structure):
    _fields_ = [("min_distance", C.c_double), ("max_distance", C.c_double), ("mean", C.c_double),
                ("std_dev", C.c_double), ("threshold", C.c_double), ("rmse", C.c_double), ("sum_sq", C.c_double),
                ("problem_count", C.c_int64), ("valid_count", C.c_int64), ("outlier_count", C.c_int64)]


class OrcParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("octree_max_points", C.c_int32), ("octree_max_depth", C.c_int32),
                ("variant", C.c_int32), ("tolerance", C.c_double), ("sigma_multiplier", C.c_double)]


class OrcIter(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("valid_points", C.c_int32), ("outlier_points", C.c_int32),
                ("has_angles", C.c_int32), ("rmse", C.c_double), ("transform", C.c_double * 16),
                ("rotation_angle", C.c_double), ("translation_distance", C.c_double)]


class OrcResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("success", C.c_int32), ("total_iterations", C.c_int32),
                ("loop_iterations", C.c_int32), ("history_len", C.c_int32), ("pad_", C.c_int32),
                ("final_rmse", C.c_double), ("final_R", C.c_double * 9), ("final_t", C.c_double * 3),
                ("last_T", C.c_double * 16), ("cum_T", C.c_double * 16)]


class OrcTrace(C.Structure):
    _fields_ = [("trace_iters", C.c_int32), ("idx", _ip), ("dist", _dp), ("mask", _u8p),
                ("stats", C.POINTER(OrcStats)), ("src_before", _dp)]


class IterRecord:
    def __init__(self, it):
        self.iteration = int(it.iteration)
        self.valid_points = int(it.valid_points if hasattr(it, "valid_points") else it.validPoints)
        self.outlier_points = int(it.outlier_points if hasattr(it, "outlier_points") else it.outlierPoints)
        self.has_angles = bool(it.has_angles)
        self.rmse = float(it.rmse)
        self.transform = np.array(list(it.transform), dtype=np.float64).reshape(4, 4)
        self.rotation_angle = float(it.rotation_angle if hasattr(it, "rotation_angle") else it.rotationAngle)
        self.translation_distance = float(
            it.translation_distance if hasattr(it, "translation_distance") else it.translationDistance)


class RunResult:
    """Uniform view of one full registration, whichever implementation produced it."""

    def __init__(self):
        self.status = 0
        self.success = False
        self.total_iterations = 0
        self.loop_iterations = 0
        self.final_rmse = 0.0
        self.final_R = np.eye(3)
        self.final_t = np.zeros(3)
        self.history: list[IterRecord] = []
        self.source_out = None
        self.trace = None
        self.message = ""
        self.signal_order = ""


class Oracle:
    """liboracle.so"""

    def __init__(self, path: str | None = None):
        self.lib = C.CDLL(path or build_oracle())
        L = self.lib
        L.orc_octree_build.restype = C.c_void_p
        L.orc_octree_build.argtypes = [_dp, C.c_int64, C.c_int, C.c_int]
        L.orc_octree_free.argtypes = [C.c_void_p]
        L.orc_octree_find_nearest.argtypes = [C.c_void_p, _dp, C.c_int64, _ip, C.c_int, C.c_int]
        L.orc_octree_count_work.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int, _i64p, _i64p, _i64p]
        L.orc_octree_dump.restype = C.c_int64
        L.orc_octree_dump.argtypes = [C.c_void_p, _i64p, _ip, _u64p, _u8p, _ip, _dp, _ip]
        L.orc_svd3.argtypes = [_dp, _dp, _dp, _dp]
        L.orc_solve_from_H.argtypes = [_dp, _dp, _dp, _dp]
        L.orc_centroids_H.argtypes = [_dp, _dp, C.c_int64, _dp, _dp, _dp]
        L.orc_kabsch.argtypes = [_dp, _dp, C.c_int64, _dp]
        L.orc_apply.argtypes = [_dp, _dp, C.c_int64]
        L.orc_mat4_mul.argtypes = [_dp, _dp, _dp]
        L.orc_angles.argtypes = [_dp, _dp, _dp]
        L.orc_iteration_stats.argtypes = [_dp, C.c_int64, _dp, C.c_int64, _ip, C.c_int, C.c_double, C.c_int, _dp,
                                          _u8p, C.POINTER(OrcStats)]
        L.orc_icp_run.restype = C.c_int
        L.orc_icp_run.argtypes = [_dp, C.c_int64, _dp, C.c_int64, C.POINTER(OrcParams), C.c_int,
                                  C.POINTER(OrcResult), C.POINTER(OrcIter), C.c_int, C.POINTER(OrcTrace), C.c_int]
        L.orc_hw_threads.restype = C.c_int

    def hw_threads(self) -> int:
        return int(self.lib.orc_hw_threads())

    # -- octree ---------------------------------------------------------------------------------
    def octree(self, tgt, max_pts=10, max_depth=20):
        return _OracleTree(self, _c3(tgt), max_pts, max_depth)

    # -- small dense pieces ---------------------------------------------------------------------
    def svd3(self, H):
        H = np.ascontiguousarray(H, dtype=np.float64).reshape(9)
        U = np.empty(9); S = np.empty(3); V = np.empty(9)
        self.lib.orc_svd3(_d(H), _d(U), _d(S), _d(V))
        return U.reshape(3, 3), S, V.reshape(3, 3)

    def solve_from_H(self, H, cA, cB):
        H = np.ascontiguousarray(H, dtype=np.float64).reshape(9)
        cA = np.ascontiguousarray(cA, dtype=np.float64); cB = np.ascontiguousarray(cB, dtype=np.float64)
        T = np.empty(16)
        self.lib.orc_solve_from_H(_d(H), _d(cA), _d(cB), _d(T))
        return T.reshape(4, 4)

    def centroids_H(self, a, b):
        a = _c3(a); b = _c3(b)
        cA = np.empty(3); cB = np.empty(3); H = np.empty(9)
        self.lib.orc_centroids_H(_d(a), _d(b), len(a), _d(cA), _d(cB), _d(H))
        return cA, cB, H.reshape(3, 3)

    def kabsch(self, a, b):
        a = _c3(a); b = _c3(b)
        T = np.empty(16)
        self.lib.orc_kabsch(_d(a), _d(b), len(a), _d(T))
        return T.reshape(4, 4)

    def apply(self, T, xyz):
        out = _c3(xyz).copy()
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        self.lib.orc_apply(_d(T), _d(out), len(out))
        return out

    def mat4_mul(self, A, B):
        A = np.ascontiguousarray(A, dtype=np.float64).reshape(16)
        B = np.ascontiguousarray(B, dtype=np.float64).reshape(16)
        Cm = np.empty(16)
        self.lib.orc_mat4_mul(_d(A), _d(B), _d(Cm))
        return Cm.reshape(4, 4)

    def angles(self, T):
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        a = C.c_double(); t = C.c_double()
        self.lib.orc_angles(_d(T), C.byref(a), C.byref(t))
        return a.value, t.value

    def iteration_stats(self, src, tgt, idx, it=0, sigma=3.0, variant=VARIANT_ENGINE):
        src = _c3(src); tgt = _c3(tgt)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        dist = np.empty(len(src)); mask = np.empty(len(src), dtype=np.uint8)
        st = OrcStats()
        self.lib.orc_iteration_stats(_d(src), len(src), _d(tgt), len(tgt), _i(idx), it, sigma, variant, _d(dist),
                                     mask.ctypes.data_as(_u8p), C.byref(st))
        return dist, mask, st

    # -- whole loop -----------------------------------------------------------------------------
    def icp(self, src, tgt, max_iterations=50, tolerance=1e-6, sigma=3.0, leaf=10, depth=20,
            variant=VARIANT_ENGINE, stop_after=-1, trace_iters=0, nthreads=1) -> RunResult:
        src = _c3(src).copy() if src is not None else None
        tgt = _c3(tgt) if tgt is not None else None
        n = 0 if src is None else len(src)
        m = 0 if tgt is None else len(tgt)
        p = OrcParams(max_iterations, leaf, depth, variant, tolerance, sigma)
        res = OrcResult()
        cap = max_iterations + 2
        hist = (OrcIter * cap)()
        tr = None
        trace = None
        if trace_iters > 0 and n > 0:
            trace = {
                "idx": np.full((trace_iters, n), -1, dtype=np.int32),
                "dist": np.zeros((trace_iters, n)),
                "mask": np.zeros((trace_iters, n), dtype=np.uint8),
                "stats": (OrcStats * trace_iters)(),
                "src_before": np.zeros((trace_iters, n, 3)),
            }
            tr = OrcTrace(trace_iters, _i(trace["idx"]), _d(trace["dist"]), trace["mask"].ctypes.data_as(_u8p),
                          trace["stats"], _d(trace["src_before"]))
        self.lib.orc_icp_run(_d(src), n, _d(tgt), m, C.byref(p), stop_after, C.byref(res), hist, cap,
                             C.byref(tr) if tr is not None else None, nthreads)
        out = RunResult()
        out.status = int(res.status)
        out.success = bool(res.success)
        out.total_iterations = int(res.total_iterations)
        out.loop_iterations = int(res.loop_iterations)
        out.final_rmse = float(res.final_rmse)
        out.final_R = np.array(list(res.final_R)).reshape(3, 3)
        out.final_t = np.array(list(res.final_t))
        out.history = [IterRecord(hist[k]) for k in range(min(res.history_len, cap))]
        out.source_out = src
        out.trace = trace
        out.last_T = np.array(list(res.last_T)).reshape(4, 4)
        out.cum_T = np.array(list(res.cum_T)).reshape(4, 4)
        return out


class _TreeDumpMixin:
    def _dump(self, fn, handle):
        n_idx = C.c_int64(0)
        n_nodes = fn(handle, C.byref(n_idx), None, None, None, None, None, None)
        depth = np.empty(n_nodes, dtype=np.int32); key = np.empty(n_nodes, dtype=np.uint64)
        leaf = np.empty(n_nodes, dtype=np.uint8); count = np.empty(n_nodes, dtype=np.int32)
        box = np.empty((n_nodes, 6)); idx = np.empty(n_idx.value, dtype=np.int32)
        fn(handle, C.byref(n_idx), _i(depth), key.ctypes.data_as(_u64p), leaf.ctypes.data_as(_u8p), _i(count),
           _d(box), _i(idx))
        return {"depth": depth, "key": key, "leaf": leaf, "count": count, "box": box, "idx": idx}


class _OracleTree(_TreeDumpMixin):
    def __init__(self, orc: Oracle, tgt, max_pts, max_depth):
        self.orc = orc
        self.tgt = tgt  # keep alive: the tree borrows the pointer
        self.h = orc.lib.orc_octree_build(_d(tgt), len(tgt), max_pts, max_depth)

    def find_nearest(self, q, variant=VARIANT_ENGINE, nthreads=1):
        q = _c3(q)
        out = np.empty(len(q), dtype=np.int32)
        self.orc.lib.orc_octree_find_nearest(self.h, _d(q), len(q), _i(out), variant, nthreads)
        return out

    def count_work(self, q, variant=VARIANT_ENGINE):
        q = _c3(q)
        a = C.c_int64(); b = C.c_int64(); c = C.c_int64()
        self.orc.lib.orc_octree_count_work(self.h, _d(q), len(q), variant, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def dump(self):
        return self._dump(self.orc.lib.orc_octree_dump, self.h)

    def close(self):
        if self.h:
            self.orc.lib.orc_octree_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


# ---------------------------------------------------------------------------------------------------------
# The compiled, unmodified reference (present only when oracle/_ref was built in the build container).
# ---------------------------------------------------------------------------------------------------------
class RefIter(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("validPoints", C.c_int32), ("outlierPoints", C.c_int32),
                ("has_angles", C.c_int32), ("rmse", C.c_double), ("transform", C.c_double * 16),
                ("rotationAngle", C.c_double), ("translationDistance", C.c_double)]


class RefResult(C.Structure):
    _fields_ = [("success", C.c_int32), ("totalIterations", C.c_int32), ("got_finished", C.c_int32),
                ("finished_ok", C.c_int32), ("n_iter_signals", C.c_int32), ("n_logs", C.c_int32),
                ("finalRMSE", C.c_double), ("finalR", C.c_double * 9), ("finalT", C.c_double * 3),
                ("finished_msg", C.c_char * 128), ("signal_order", C.c_char * 4096)]


def ref_available() -> bool:
    return os.path.exists(REF_ENGINE_SO) and os.path.exists(REF_CLI_SO)


class RefEngine(_TreeDumpMixin):
    """oracle/_ref/libref_engine.so: PointCloudRegistration/core compiled unmodified."""

    def __init__(self):
        self.lib = C.CDLL(REF_ENGINE_SO)
        L = self.lib
        L.ref_octree_create.restype = C.c_void_p
        L.ref_octree_create.argtypes = [_dp, C.c_int64, C.c_int, C.c_int]
        L.ref_octree_destroy.argtypes = [C.c_void_p]
        L.ref_octree_find_nearest.argtypes = [C.c_void_p, _dp, C.c_int64, _ip, C.c_int]
        L.ref_octree_dump.restype = C.c_int64
        L.ref_octree_dump.argtypes = [C.c_void_p, _i64p, _ip, _u64p, _u8p, _ip, _dp, _ip]
        L.ref_engine_run.restype = C.c_int
        L.ref_engine_run.argtypes = [_dp, C.c_int64, _dp, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_int,
                                     C.c_int, C.c_int, C.POINTER(RefResult), C.POINTER(RefIter), C.c_int, C.c_int]
        L.ref_engine_kabsch.argtypes = [_dp, _dp, C.c_int64, _dp]
        L.ref_engine_centroids_H.argtypes = [_dp, _dp, C.c_int64, _dp, _dp, _dp]
        L.ref_engine_solve_from_H.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.ref_engine_apply.argtypes = [_dp, _dp, C.c_int64]
        L.ref_engine_mat4_mul.argtypes = [_dp, _dp, _dp]
        L.ref_engine_angles.argtypes = [_dp, _dp, _dp]
        L.ref_max_threads.restype = C.c_int

    def max_threads(self):
        return int(self.lib.ref_max_threads())

    def octree(self, tgt, max_pts=10, max_depth=20):
        return _RefTree(self, _c3(tgt), max_pts, max_depth)

    def solve_from_H(self, H, cA, cB):
        H = np.ascontiguousarray(H, dtype=np.float64).reshape(9)
        cA = np.ascontiguousarray(cA, dtype=np.float64); cB = np.ascontiguousarray(cB, dtype=np.float64)
        T = np.empty(16); U = np.empty(9); S = np.empty(3); V = np.empty(9)
        self.lib.ref_engine_solve_from_H(_d(H), _d(cA), _d(cB), _d(T), _d(U), _d(S), _d(V))
        return T.reshape(4, 4), U.reshape(3, 3), S, V.reshape(3, 3)

    def centroids_H(self, a, b):
        a = _c3(a); b = _c3(b)
        cA = np.empty(3); cB = np.empty(3); H = np.empty(9)
        self.lib.ref_engine_centroids_H(_d(a), _d(b), len(a), _d(cA), _d(cB), _d(H))
        return cA, cB, H.reshape(3, 3)

    def kabsch(self, a, b):
        a = _c3(a); b = _c3(b)
        T = np.empty(16)
        self.lib.ref_engine_kabsch(_d(a), _d(b), len(a), _d(T))
        return T.reshape(4, 4)

    def apply(self, T, xyz):
        out = _c3(xyz).copy()
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        self.lib.ref_engine_apply(_d(T), _d(out), len(out))
        return out

    def mat4_mul(self, A, B):
        A = np.ascontiguousarray(A, dtype=np.float64).reshape(16)
        B = np.ascontiguousarray(B, dtype=np.float64).reshape(16)
        Cm = np.empty(16)
        self.lib.ref_engine_mat4_mul(_d(A), _d(B), _d(Cm))
        return Cm.reshape(4, 4)

    def angles(self, T):
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        a = C.c_double(); t = C.c_double()
        self.lib.ref_engine_angles(_d(T), C.byref(a), C.byref(t))
        return a.value, t.value

    def last_logs(self) -> list:
        """logMessage texts of the last run (core/icpengine.h:75), in order."""
        if not hasattr(self.lib, "ref_engine_last_logs"):
            return []
        self.lib.ref_engine_last_logs.restype = C.c_int64
        self.lib.ref_engine_last_logs.argtypes = [C.c_char_p, C.c_int64]
        need = self.lib.ref_engine_last_logs(None, 0)
        buf = C.create_string_buffer(int(need) + 1)
        self.lib.ref_engine_last_logs(buf, need + 1)
        return [ln for ln in buf.value.decode("utf-8", "replace").split("\n") if ln != ""]

    def icp(self, src, tgt, max_iterations=50, tolerance=1e-6, sigma=3.0, leaf=10, depth=20, stop_after=-1,
            print_logs=False) -> RunResult:
        src = _c3(src).copy() if src is not None else None
        tgt = _c3(tgt) if tgt is not None else None
        n = 0 if src is None else len(src)
        m = 0 if tgt is None else len(tgt)
        res = RefResult()
        cap = max_iterations + 2
        hist = (RefIter * cap)()
        rc = self.lib.ref_engine_run(_d(src), n, _d(tgt), m, max_iterations, tolerance, sigma, leaf, depth,
                                     stop_after, C.byref(res), hist, cap, 1 if print_logs else 0)
        out = RunResult()
        out.logs = self.last_logs()
        out.message = res.finished_msg.decode("utf-8", "replace")
        out.signal_order = res.signal_order.decode()
        out.success = bool(res.success)
        out.total_iterations = int(res.totalIterations)
        out.final_rmse = float(res.finalRMSE)
        out.final_R = np.array(list(res.finalR)).reshape(3, 3)
        out.final_t = np.array(list(res.finalT))
        nh = out.signal_order.count("i")
        out.history = [IterRecord(hist[k]) for k in range(min(nh, cap))] if rc == 0 else []
        out.source_out = src
        msg = out.message
        if rc != 0 or not res.got_finished:
            out.status = 1
        elif res.finished_ok:
            out.status = 0
        elif "取消" in msg:      # 用户取消
            out.status = 2
        elif "不足" in msg:      # 有效点对不足
            out.status = 3
        else:
            out.status = 1               # 点云数据为空 / 源点云或目标点云为空
        return out


class _RefTree(_TreeDumpMixin):
    def __init__(self, ref: RefEngine, tgt, max_pts, max_depth):
        self.ref = ref
        self.h = ref.lib.ref_octree_create(_d(tgt), len(tgt), max_pts, max_depth)

    def find_nearest(self, q, nthreads=1):
        q = _c3(q)
        out = np.empty(len(q), dtype=np.int32)
        self.ref.lib.ref_octree_find_nearest(self.h, _d(q), len(q), _i(out), nthreads)
        return out

    def dump(self):
        return self._dump(self.ref.lib.ref_octree_dump, self.h)

    def close(self):
        if self.h:
            self.ref.lib.ref_octree_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class RefCli:
    """oracle/_ref/libref_cli.so: icp_registration.cpp compiled unmodified."""

    def __init__(self):
        self.lib = C.CDLL(REF_CLI_SO)
        L = self.lib
        L.ref_cli_octree_create.restype = C.c_void_p
        L.ref_cli_octree_create.argtypes = [_dp, C.c_int64, C.c_int, C.c_int]
        L.ref_cli_octree_destroy.argtypes = [C.c_void_p]
        L.ref_cli_octree_find_nearest.argtypes = [C.c_void_p, _dp, C.c_int64, _ip]
        L.ref_cli_icp.restype = C.c_int
        L.ref_cli_icp.argtypes = [_dp, C.c_int64, _dp, C.c_int64, C.c_int, C.c_double, _dp, _dp, _dp, C.c_int, C.c_int]
        L.ref_cli_best_fit_transform.argtypes = [_dp, _dp, C.c_int64, _dp]
        L.ref_cli_save_transformation.argtypes = [_dp, _dp, _dp, C.c_int, C.c_char_p]

    def find_nearest(self, tgt, q, max_pts=10, max_depth=20):
        tgt = _c3(tgt); q = _c3(q)
        h = self.lib.ref_cli_octree_create(_d(tgt), len(tgt), max_pts, max_depth)
        out = np.empty(len(q), dtype=np.int32)
        self.lib.ref_cli_octree_find_nearest(h, _d(q), len(q), _i(out))
        self.lib.ref_cli_octree_destroy(h)
        return out

    def icp(self, src, tgt, max_iterations=20, tolerance=1e-2):
        src = _c3(src).copy(); tgt = _c3(tgt)
        R = np.empty(9); t = np.empty(3)
        its = np.zeros((max_iterations + 1, 16))
        n = self.lib.ref_cli_icp(_d(src), len(src), _d(tgt), len(tgt), max_iterations, tolerance, _d(R), _d(t),
                                 _d(its), max_iterations + 1, 0)
        return src, R.reshape(3, 3), t, its[:n].reshape(n, 4, 4)

    def best_fit_transform(self, a, b):
        a = _c3(a); b = _c3(b)
        T = np.empty(16)
        self.lib.ref_cli_best_fit_transform(_d(a), _d(b), len(a), _d(T))
        return T.reshape(4, 4)

    def save_transformation(self, R, t, its, filename: str):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
        t = np.ascontiguousarray(t, dtype=np.float64)
        its = np.ascontiguousarray(its, dtype=np.float64).reshape(-1, 16)
        self.lib.ref_cli_save_transformation(_d(R), _d(t), _d(its), len(its), filename.encode())


# ---------------------------------------------------------------------------------------------------------
# Data-format steps either side of the loop: LAS decode / encode, bounds, downsampling,
# replay, transformation report.  OracleIO = cloudio_oracle.c (this repo's restatement); RefIO = the unmodified
# reference (oracle/_ref/libref_io.so + the CLI's functions in libref_cli.so).
# ---------------------------------------------------------------------------------------------------------
REF_IO_SO = os.path.join(HERE, "_ref", "libref_io.so")
LAS_HEADER_BYTES = 227
LAS_RECORD_BYTES = 20


def _u8(a):
    return a.ctypes.data_as(_u8p)


class OracleIO:
    def __init__(self, path: str | None = None):
        self.lib = C.CDLL(path or build_oracle())
        L = self.lib
        L.orc_las_decode.argtypes = [_u8p, C.c_int64, C.c_int32, _dp, _dp, _dp]
        L.orc_las_encode.argtypes = [_dp, C.c_int64, _dp, _dp, _u8p]
        L.orc_bounds.argtypes = [_dp, C.c_int64, _dp, _dp]
        L.orc_las_file_image.restype = C.c_int64
        L.orc_las_file_image.argtypes = [C.c_int, _dp, C.c_int64, _dp, _dp, _u8p]
        L.orc_las_parse_header.restype = C.c_int
        L.orc_las_parse_header.argtypes = [_u8p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), _dp, _dp]
        L.orc_downsample.restype = C.c_int64
        L.orc_downsample.argtypes = [_dp, C.c_int64, C.c_int32, _dp]
        L.orc_downsample_stride.restype = C.c_int64
        L.orc_downsample_stride.argtypes = [_dp, C.c_int64, C.c_int64, _dp]
        L.orc_cloud_apply.argtypes = [_dp, _dp, C.c_int64, _dp]
        L.orc_transformation_text.restype = C.c_int64
        L.orc_transformation_text.argtypes = [_dp, _dp, _dp, C.c_int32, C.c_char_p, C.c_int64]

    def las_decode(self, records, n, record_length, scale, offset):
        records = np.ascontiguousarray(records, dtype=np.uint8)
        scale = np.ascontiguousarray(scale, dtype=np.float64); offset = np.ascontiguousarray(offset, dtype=np.float64)
        out = np.empty((n, 3))
        self.lib.orc_las_decode(_u8(records), n, record_length, _d(scale), _d(offset), _d(out))
        return out

    def las_encode(self, xyz, scale, offset):
        xyz = _c3(xyz)
        scale = np.ascontiguousarray(scale, dtype=np.float64); offset = np.ascontiguousarray(offset, dtype=np.float64)
        out = np.empty(len(xyz) * LAS_RECORD_BYTES, dtype=np.uint8)
        self.lib.orc_las_encode(_d(xyz), len(xyz), _d(scale), _d(offset), _u8(out))
        return out

    def bounds(self, xyz):
        xyz = _c3(xyz)
        mn = np.empty(3); mx = np.empty(3)
        self.lib.orc_bounds(_d(xyz), len(xyz), _d(mn), _d(mx))
        return mn, mx

    def las_file_image(self, xyz, variant=VARIANT_ENGINE, scale=(0.001,) * 3, offset=(0.0,) * 3):
        xyz = _c3(xyz)
        scale = np.ascontiguousarray(scale, dtype=np.float64); offset = np.ascontiguousarray(offset, dtype=np.float64)
        out = np.empty(LAS_HEADER_BYTES + LAS_RECORD_BYTES * len(xyz), dtype=np.uint8)
        n = self.lib.orc_las_file_image(variant, _d(xyz), len(xyz), _d(scale), _d(offset), _u8(out))
        return out[:n]

    def las_parse_header(self, header):
        header = np.ascontiguousarray(header, dtype=np.uint8)
        off = C.c_uint32(); n = C.c_uint32(); rl = C.c_uint16()
        scale = np.empty(3); offset = np.empty(3)
        ok = self.lib.orc_las_parse_header(_u8(header), C.byref(off), C.byref(n), C.byref(rl), _d(scale), _d(offset))
        return bool(ok), off.value, n.value, rl.value, scale, offset

    def las_read_image(self, image, max_points=0):
        """What LASIO::readLAS returns for a file with these bytes."""
        image = np.ascontiguousarray(image, dtype=np.uint8)
        ok, off, n, rl, scale, offset = self.las_parse_header(image[:LAS_HEADER_BYTES])
        if not ok:
            return None
        if max_points and max_points < n:
            n = max_points
        return self.las_decode(image[off:off + n * rl], n, rl, scale, offset)

    def downsample(self, xyz, target):
        xyz = _c3(xyz)
        out = np.empty((max(min(len(xyz), max(target, 0)), 1), 3))
        n = self.lib.orc_downsample(_d(xyz), len(xyz), target, _d(out))
        return out[:n]

    def downsample_stride(self, xyz, stride):
        xyz = _c3(xyz)
        out = np.empty((len(xyz) // max(stride, 1) + 1, 3))
        n = self.lib.orc_downsample_stride(_d(xyz), len(xyz), stride, _d(out))
        return out[:n]

    def cloud_apply(self, T, xyz):
        xyz = _c3(xyz)
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        out = np.empty_like(xyz)
        self.lib.orc_cloud_apply(_d(T), _d(xyz), len(xyz), _d(out))
        return out

    def transformation_text(self, R, t, its=None) -> bytes:
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
        t = np.ascontiguousarray(t, dtype=np.float64)
        its = np.zeros((0, 16)) if its is None else np.ascontiguousarray(its, dtype=np.float64).reshape(-1, 16)
        need = self.lib.orc_transformation_text(_d(R), _d(t), _d(its) if len(its) else None, len(its), None, 0)
        buf = C.create_string_buffer(need + 1)
        self.lib.orc_transformation_text(_d(R), _d(t), _d(its) if len(its) else None, len(its), buf, need + 1)
        return buf.raw[:need]


def ref_io_available() -> bool:
    return os.path.exists(REF_IO_SO) and os.path.exists(REF_CLI_SO)


class RefIO:
    """The unmodified reference: LASIO / PointCloud (engine) and readLASFile / saveResultAsLAS / saveTransformation (CLI)."""

    def __init__(self):
        self.lib = C.CDLL(REF_IO_SO)
        self.cli = C.CDLL(REF_CLI_SO)
        L = self.lib
        L.ref_io_write_las.restype = C.c_int
        L.ref_io_write_las.argtypes = [C.c_char_p, _dp, C.c_int64]
        L.ref_io_read_las.restype = C.c_int64
        L.ref_io_read_las.argtypes = [C.c_char_p, C.c_int64, _dp, C.c_int64, _dp]
        L.ref_io_read_las_batch.restype = C.c_int64
        L.ref_io_read_las_batch.argtypes = [C.c_char_p, C.c_int64, _dp, C.c_int64, _i64p, C.c_int64, _i64p]
        L.ref_io_bounds.argtypes = [_dp, C.c_int64, _dp, _dp]
        L.ref_io_downsample.restype = C.c_int64
        L.ref_io_downsample.argtypes = [_dp, C.c_int64, C.c_int, _dp, C.c_int64]
        L.ref_io_apply_transform.argtypes = [_dp, _dp, _dp, C.c_int64]
        K = self.cli
        K.ref_cli_read_las.restype = C.c_int64
        K.ref_cli_read_las.argtypes = [C.c_char_p, _dp, C.c_int64, _dp, _dp]
        K.ref_cli_save_las.argtypes = [_dp, C.c_int64, _dp, _dp, C.c_char_p]
        K.ref_cli_save_transformation.argtypes = [_dp, _dp, _dp, C.c_int, C.c_char_p]

    def write_las(self, path, xyz) -> bool:
        xyz = _c3(xyz)
        return bool(self.lib.ref_io_write_las(path.encode(), _d(xyz), len(xyz)))

    def read_las(self, path, max_points=0):
        b = np.empty(6)
        n = self.lib.ref_io_read_las(path.encode(), max_points, None, 0, _d(b))
        if n < 0:
            return None, None
        out = np.empty((n, 3))
        self.lib.ref_io_read_las(path.encode(), max_points, _d(out), n, _d(b))
        return out, b

    def read_las_batch(self, path, batch_size, cap):
        out = np.empty((cap, 3)); sizes = np.zeros(cap + 1, dtype=np.int64); nb = C.c_int64()
        n = self.lib.ref_io_read_las_batch(path.encode(), batch_size, _d(out), cap, sizes.ctypes.data_as(_i64p), len(sizes), C.byref(nb))
        return out[:n], sizes[:nb.value]

    def bounds(self, xyz):
        xyz = _c3(xyz)
        mn = np.empty(3); mx = np.empty(3)
        self.lib.ref_io_bounds(_d(xyz), len(xyz), _d(mn), _d(mx))
        return mn, mx

    def downsample(self, xyz, target):
        xyz = _c3(xyz)
        out = np.empty((max(len(xyz), 1), 3))
        n = self.lib.ref_io_downsample(_d(xyz), len(xyz), target, _d(out), len(out))
        return None if n < 0 else out[:n]

    def apply_transform(self, R, t, xyz):
        xyz = _c3(xyz).copy()
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(9); t = np.ascontiguousarray(t, dtype=np.float64)
        self.lib.ref_io_apply_transform(_d(R), _d(t), _d(xyz), len(xyz))
        return xyz

    def cli_read_las(self, path):
        s = np.empty(3); o = np.empty(3)
        n = self.cli.ref_cli_read_las(path.encode(), None, 0, _d(s), _d(o))
        if n < 0:
            return None, None, None
        out = np.empty((n, 3))
        self.cli.ref_cli_read_las(path.encode(), _d(out), n, _d(s), _d(o))
        return out, s, o

    def cli_save_las(self, path, xyz, scale, offset):
        xyz = _c3(xyz)
        scale = np.ascontiguousarray(scale, dtype=np.float64); offset = np.ascontiguousarray(offset, dtype=np.float64)
        self.cli.ref_cli_save_las(_d(xyz), len(xyz), _d(scale), _d(offset), path.encode())

    def cli_save_transformation(self, path, R, t, its=None):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(9); t = np.ascontiguousarray(t, dtype=np.float64)
        its = np.zeros((0, 16)) if its is None else np.ascontiguousarray(its, dtype=np.float64).reshape(-1, 16)
        self.cli.ref_cli_save_transformation(_d(R), _d(t), _d(its) if len(its) else None, len(its), path.encode())
