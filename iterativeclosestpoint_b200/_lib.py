"""ctypes binding of the C ABI in include/icp_b200.h (libicp_b200.so, built in-tree by __graft_entry__.build()).

There is no CPU fallback: if the shared library is missing, or the process has no CUDA device, every entry
point raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libicp_b200.so")

ICP_OK = 0
ICP_EMPTY_INPUT = 1
ICP_CANCELLED = 2
ICP_TOO_FEW_INLIERS = 3
ICP_INVALID_ARGUMENT = 4
ICP_CUDA_ERROR = 5
ICP_NCCL_ERROR = 6
ICP_NO_OCTREE = 7
ICP_IO_ERROR = 8
ICP_BAD_FORMAT = 9
STATUS_NAMES = {0: "OK", 1: "EMPTY_INPUT", 2: "CANCELLED", 3: "TOO_FEW_INLIERS", 4: "INVALID_ARGUMENT",
                5: "CUDA_ERROR", 6: "NCCL_ERROR", 7: "NO_OCTREE", 8: "IO_ERROR", 9: "BAD_FORMAT"}
LAS_HEADER_BYTES = 227
LAS_RECORD_BYTES = 20

VARIANT_ENGINE = 0
VARIANT_CLI = 1

# every symbol include/icp_b200.h declares (tests check the library exports exactly these)
EXPORTED = [
    "icp_create", "icp_destroy", "icp_last_error", "icp_abi_version", "icp_set_params", "icp_get_params",
    "icp_default_params", "icp_set_callbacks", "icp_set_option", "icp_register", "icp_source_upload",
    "icp_register_resident", "icp_octree_build", "icp_octree_get_info", "icp_octree_dump", "icp_nn_query",
    "icp_iteration_stats", "icp_best_fit_transform", "icp_solve_from_H", "icp_apply_transform",
    "icp_comm_unique_id", "icp_comm_init", "icp_comm_destroy", "icp_register_sharded", "icp_register_batch",
    "icp_kernel_launches", "icp_nn_counters", "icp_nn_tile_counters",
    "icp_las_parse_header", "icp_las_decode", "icp_las_encode", "icp_cloud_bounds", "icp_las_file_image", "icp_las_write",
    "icp_las_read", "icp_downsample", "icp_downsample_stride", "icp_replay_iteration", "icp_save_transformation",
    "icp_register_las",
]


class IcpParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("octree_max_points", C.c_int32), ("octree_max_depth", C.c_int32),
                ("variant", C.c_int32), ("tolerance", C.c_double), ("sigma_multiplier", C.c_double)]


class IcpIteration(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("valid_points", C.c_int32), ("outlier_points", C.c_int32),
                ("has_angles", C.c_int32), ("rmse", C.c_double), ("transform", C.c_double * 16),
                ("rotation_angle", C.c_double), ("translation_distance", C.c_double),
                ("nn_ms", C.c_double), ("iter_ms", C.c_double)]


class IcpStats(C.Structure):
    _fields_ = [("min_distance", C.c_double), ("max_distance", C.c_double), ("mean", C.c_double),
                ("std_dev", C.c_double), ("threshold", C.c_double), ("rmse", C.c_double), ("sum_sq", C.c_double),
                ("problem_count", C.c_int64), ("valid_count", C.c_int64), ("outlier_count", C.c_int64)]


class IcpResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("success", C.c_int32), ("total_iterations", C.c_int32),
                ("loop_iterations", C.c_int32), ("history_len", C.c_int32), ("history_cap", C.c_int32),
                ("final_rmse", C.c_double), ("final_R", C.c_double * 9), ("final_t", C.c_double * 3),
                ("cumulative_T", C.c_double * 16), ("last_T", C.c_double * 16),
                ("history", C.POINTER(IcpIteration)),
                ("ms_h2d", C.c_float), ("ms_build", C.c_float), ("ms_loop", C.c_float), ("ms_d2h", C.c_float),
                ("ms_nn_total", C.c_float), ("ms_nn_first", C.c_float)]


class IcpOctreeInfo(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("n_nodes", C.c_int64), ("n_leaves", C.c_int64), ("node_bytes", C.c_int64),
                ("point_bytes", C.c_int64), ("depth", C.c_int32), ("max_points", C.c_int32), ("max_depth", C.c_int32),
                ("pad_", C.c_int32), ("root_lo", C.c_double * 3), ("root_hi", C.c_double * 3),
                ("build_ms", C.c_float), ("pad2_", C.c_float),
                ("search_nodes", C.c_int64), ("search_node_bytes", C.c_int64), ("grid_bytes", C.c_int64),
                ("search_depth", C.c_int32), ("grid_base_level", C.c_int32), ("grid_fine_level", C.c_int32),
                ("pad3_", C.c_int32), ("grid_base_cell", C.c_double)]


class IcpLasHeader(C.Structure):
    _fields_ = [("offset_to_data", C.c_uint32), ("n_points", C.c_uint32), ("record_length", C.c_uint16),
                ("pad_", C.c_uint16 * 3), ("scale", C.c_double * 3), ("offset", C.c_double * 3),
                ("min", C.c_double * 3), ("max", C.c_double * 3)]


class IcpLasPoints(C.Structure):
    _fields_ = [("records", C.c_void_p), ("n", C.c_int64), ("record_length", C.c_int32), ("pad_", C.c_int32),
                ("scale", C.c_double * 3), ("offset", C.c_double * 3)]


ITERATION_CB = C.CFUNCTYPE(None, C.POINTER(IcpIteration), C.c_void_p)
PROGRESS_CB = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_double, C.c_void_p)
LOG_CB = C.CFUNCTYPE(None, C.c_char_p, C.c_void_p)

_dp = C.POINTER(C.c_double)
_lib = None


class IcpError(RuntimeError):
    def __init__(self, status: int, detail: str = ""):
        self.status = status
        super().__init__(f"icp_b200: {STATUS_NAMES.get(status, status)}" + (f": {detail}" if detail else ""))


def load() -> C.CDLL:
    """Loads libicp_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C iterativeclosestpoint_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.icp_create.argtypes = [C.POINTER(vp), C.c_int]
    L.icp_destroy.argtypes = [vp]
    L.icp_destroy.restype = None
    L.icp_last_error.argtypes = [vp]
    L.icp_last_error.restype = C.c_char_p
    L.icp_set_params.argtypes = [vp, C.POINTER(IcpParams)]
    L.icp_get_params.argtypes = [vp, C.POINTER(IcpParams)]
    L.icp_default_params.argtypes = [C.POINTER(IcpParams)]
    L.icp_default_params.restype = None
    L.icp_set_callbacks.argtypes = [vp, ITERATION_CB, PROGRESS_CB, LOG_CB, vp]
    L.icp_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.icp_register.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, C.POINTER(IcpResult), vp]
    L.icp_source_upload.argtypes = [vp, vp, C.c_int64]
    L.icp_register_resident.argtypes = [vp, C.c_int64, C.POINTER(IcpResult), vp, vp]
    L.icp_octree_build.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_int]
    L.icp_octree_get_info.argtypes = [vp, C.POINTER(IcpOctreeInfo)]
    L.icp_octree_dump.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp, vp, vp, vp, vp, vp]
    L.icp_nn_query.argtypes = [vp, vp, C.c_int64, vp, vp, C.POINTER(C.c_float)]
    L.icp_iteration_stats.argtypes = [vp, vp, C.c_int64, vp, C.c_int, vp, vp, C.POINTER(IcpStats)]
    L.icp_best_fit_transform.argtypes = [vp, vp, vp, C.c_int64, vp]
    L.icp_solve_from_H.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.icp_apply_transform.argtypes = [vp, vp, vp, C.c_int64]
    L.icp_comm_unique_id.argtypes = [vp, vp]
    L.icp_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.icp_comm_destroy.argtypes = [vp]
    L.icp_register_sharded.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, C.POINTER(IcpResult), vp]
    L.icp_register_batch.argtypes = [vp, C.c_int32, vp, vp, vp, vp, C.POINTER(IcpResult)]
    L.icp_nn_counters.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]
    L.icp_nn_tile_counters.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]
    i64p = C.POINTER(C.c_int64)
    L.icp_las_parse_header.argtypes = [vp, C.POINTER(IcpLasHeader)]
    L.icp_las_decode.argtypes = [vp, vp, C.c_int64, C.c_int32, vp, vp, vp]
    L.icp_las_encode.argtypes = [vp, vp, C.c_int64, vp, vp, vp]
    L.icp_cloud_bounds.argtypes = [vp, vp, C.c_int64, vp, vp]
    L.icp_las_file_image.argtypes = [vp, vp, C.c_int64, C.c_int, vp, vp, vp, C.c_int64, i64p]
    L.icp_las_write.argtypes = [vp, C.c_char_p, vp, C.c_int64, C.c_int, vp, vp]
    L.icp_las_read.argtypes = [vp, C.c_char_p, C.c_int64, C.c_int, C.POINTER(IcpLasHeader), vp, C.c_int64, i64p]
    L.icp_downsample.argtypes = [vp, vp, C.c_int64, C.c_int32, vp, i64p]
    L.icp_downsample_stride.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, i64p]
    L.icp_replay_iteration.argtypes = [vp, vp, C.c_int64, vp, vp]
    L.icp_save_transformation.argtypes = [C.c_char_p, vp, vp, vp, C.c_int32]
    L.icp_register_las.argtypes = [vp, C.POINTER(IcpLasPoints), C.POINTER(IcpLasPoints), C.POINTER(IcpResult), vp, vp]
    L.icp_kernel_launches.argtypes = [vp]
    L.icp_kernel_launches.restype = C.c_int64
    _lib = L
    return L
