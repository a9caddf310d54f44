"""Host-side mirror of the reference's data-format interface next to the ICP loop, over the C ABI.

    LASIO.readLAS / writeLAS / readLASBatch            PointCloudRegistration/core/lasio.h:21-40
    PointCloud.points / computeBounds / applyTransform /
               applyTransformMatrix / downsample         core/pointcloud.h:30-65, pointcloud.cpp:24-128
    readLASFile / saveResultAsLAS / saveTransformation   icp_registration.cpp:248, 698, 625
    sample_stride                                        icp_registration.cpp:857,877-882
    replay                                               widgets/pointcloudviewer.cpp:86-116

Point work (decode, encode, bounds, gathers, transforms) runs in libicp_b200.so's CUDA kernels (csrc/cloudio.cu); only the
227-byte header framing and the text report are host code.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import VARIANT_CLI, VARIANT_ENGINE
from .engine import Handle, _c3, _ptr, _result_from_c

_default_handle = None


def _h(handle):
    global _default_handle
    if handle is not None:
        return handle
    if _default_handle is None:
        _default_handle = Handle(0)
    return _default_handle


def _v3(a):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(3)


# ---- stage API (arrays in, arrays out) ---------------------------------------------------------------------
def las_parse_header(header) -> _lib.IcpLasHeader:
    header = np.ascontiguousarray(header, dtype=np.uint8)
    if header.size < _lib.LAS_HEADER_BYTES:
        raise ValueError("a LAS 1.2 header has 227 bytes")
    out = _lib.IcpLasHeader()
    st = _lib.load().icp_las_parse_header(_ptr(header), C.byref(out))
    if st != 0:
        raise _lib.IcpError(st)
    return out


def las_decode(records, n, record_length, scale, offset, handle=None) -> np.ndarray:
    h = _h(handle)
    records = np.ascontiguousarray(records, dtype=np.uint8)
    assert records.size >= n * record_length
    scale = _v3(scale); offset = _v3(offset)
    out = np.empty((n, 3))
    h.check(h.lib.icp_las_decode(h.h, _ptr(records), n, record_length, _ptr(scale), _ptr(offset), _ptr(out)))
    return out


def las_encode(xyz, scale, offset, handle=None) -> np.ndarray:
    h = _h(handle)
    xyz = _c3(xyz)
    scale = _v3(scale); offset = _v3(offset)
    out = np.empty(len(xyz) * _lib.LAS_RECORD_BYTES, dtype=np.uint8)
    h.check(h.lib.icp_las_encode(h.h, _ptr(xyz), len(xyz), _ptr(scale), _ptr(offset), _ptr(out)))
    return out


def cloud_bounds(xyz, handle=None):
    h = _h(handle)
    xyz = _c3(xyz)
    mn = np.empty(3); mx = np.empty(3)
    h.check(h.lib.icp_cloud_bounds(h.h, _ptr(xyz), len(xyz), _ptr(mn), _ptr(mx)))
    return mn, mx


def las_file_image(xyz, variant=VARIANT_ENGINE, scale=(0.001,) * 3, offset=(0.0,) * 3, handle=None) -> np.ndarray:
    h = _h(handle)
    xyz = _c3(xyz)
    scale = _v3(scale); offset = _v3(offset)
    out = np.empty(_lib.LAS_HEADER_BYTES + _lib.LAS_RECORD_BYTES * len(xyz), dtype=np.uint8)
    n = C.c_int64()
    h.check(h.lib.icp_las_file_image(h.h, _ptr(xyz), len(xyz), variant, _ptr(scale), _ptr(offset), _ptr(out), out.size, C.byref(n)))
    return out[:n.value]


def downsample(xyz, target_size, handle=None):
    """PointCloud::downsample: None where the reference returns nullptr."""
    h = _h(handle)
    xyz = _c3(xyz)
    out = np.empty((max(min(len(xyz), max(int(target_size), 0)), 1), 3))
    n = C.c_int64()
    st = h.lib.icp_downsample(h.h, _ptr(xyz), len(xyz), int(target_size), _ptr(out), C.byref(n))
    if st == _lib.ICP_EMPTY_INPUT:
        return None
    h.check(st)
    return out[:n.value]


def sample_stride(xyz, sample_rate, handle=None) -> np.ndarray:
    h = _h(handle)
    xyz = _c3(xyz)
    out = np.empty((len(xyz) // int(sample_rate) + 1, 3))
    n = C.c_int64()
    h.check(h.lib.icp_downsample_stride(h.h, _ptr(xyz), len(xyz), int(sample_rate), _ptr(out), C.byref(n)))
    return out[:n.value]


def replay(original, transform=None, handle=None) -> np.ndarray:
    """The source cloud as the viewer shows it for one history record (transform=None: the original)."""
    h = _h(handle)
    original = _c3(original)
    out = np.empty_like(original)
    T = None if transform is None else np.ascontiguousarray(transform, dtype=np.float64).reshape(16)
    h.check(h.lib.icp_replay_iteration(h.h, _ptr(original), len(original), _ptr(T), _ptr(out)))
    return out


def register_las(src_records, src_header, tgt_records, tgt_header, handle=None, stop_flag=None):
    """icp_register_las: both clouds as raw LAS point records (+ their parsed headers).  Returns (ICPResult, registered source)."""
    h = _h(handle)
    p = _lib.IcpParams()
    h.lib.icp_get_params(h.h, C.byref(p))
    res, hist = h._result(p.max_iterations + 2)
    pts = []
    keep = []
    for rec, hd in ((src_records, src_header), (tgt_records, tgt_header)):
        rec = np.ascontiguousarray(rec, dtype=np.uint8)
        keep.append(rec)
        q = _lib.IcpLasPoints()
        q.records = rec.ctypes.data
        q.n = rec.size // hd.record_length
        q.record_length = hd.record_length
        for a in range(3):
            q.scale[a] = hd.scale[a]
            q.offset[a] = hd.offset[a]
        pts.append(q)
    out = np.empty((pts[0].n, 3))
    st = h.lib.icp_register_las(h.h, C.byref(pts[0]), C.byref(pts[1]), C.byref(res), _ptr(out),
                                C.byref(stop_flag) if stop_flag is not None else None)
    h.check(st, ok=(0, 1, 2, 3))
    return _result_from_c(res, hist), out


# ---- the reference's classes and free functions ------------------------------------------------------------------
class PointCloud:
    """core/pointcloud.h:30-65 without the display attributes: `points` is an (n, 3) float64 array."""

    def __init__(self, points=None, handle=None):
        self.points = np.empty((0, 3)) if points is None else _c3(points)
        self.minX = self.maxX = self.minY = self.maxY = self.minZ = self.maxZ = 0.0
        # the CLI's PointCloud carries the LAS scale / offset of the file it came from (icp_registration.cpp:213-218)
        self.x_scale = self.y_scale = self.z_scale = 0.001
        self.x_offset = self.y_offset = self.z_offset = 0.0
        self._handle = handle

    def size(self):
        return len(self.points)

    def empty(self):
        return len(self.points) == 0

    def clear(self):
        self.points = np.empty((0, 3))

    def computeBounds(self):  # pointcloud.cpp:24-45
        mn, mx = cloud_bounds(self.points, self._handle)
        self.minX, self.minY, self.minZ = (float(v) for v in mn)
        self.maxX, self.maxY, self.maxZ = (float(v) for v in mx)

    def applyTransform(self, R, t):  # pointcloud.cpp:73-86
        T = np.eye(4)
        T[:3, :3] = np.asarray(R, dtype=np.float64).reshape(3, 3)
        T[:3, 3] = np.asarray(t, dtype=np.float64).reshape(3)
        self.points = replay(self.points, T, self._handle)

    def applyTransformMatrix(self, transform):  # pointcloud.cpp:88-105: silently ignores anything but a 4x4
        T = np.asarray(transform, dtype=np.float64)
        if T.shape != (4, 4):
            return
        Tm = T.copy()
        Tm[3] = [0.0, 0.0, 0.0, 1.0]
        self.points = replay(self.points, Tm, self._handle)

    def downsample(self, targetSize):  # pointcloud.cpp:107-128
        pts = downsample(self.points, targetSize, self._handle)
        return None if pts is None else PointCloud(pts, self._handle)


class LASIO:
    """core/lasio.h:15-40."""

    @staticmethod
    def readLAS(filename: str, cloud: PointCloud, maxPoints: int = 0, handle=None) -> bool:
        h = _h(handle or cloud._handle)
        hdr = _lib.IcpLasHeader()
        n = C.c_int64()
        st = h.lib.icp_las_read(h.h, filename.encode(), int(maxPoints), VARIANT_ENGINE, C.byref(hdr), None, 0, C.byref(n))
        if st != 0:
            return False
        pts = np.empty((n.value, 3))
        st = h.lib.icp_las_read(h.h, filename.encode(), int(maxPoints), VARIANT_ENGINE, C.byref(hdr), _ptr(pts), n.value, C.byref(n))
        if st != 0:
            return False
        cloud.points = pts
        cloud.computeBounds()  # lasio.cpp:113
        return True

    @staticmethod
    def writeLAS(filename: str, cloud: PointCloud, handle=None) -> bool:
        h = _h(handle or cloud._handle)
        st = h.lib.icp_las_write(h.h, filename.encode(), _ptr(_c3(cloud.points)), len(cloud.points), VARIANT_ENGINE, None, None)
        return st == 0

    @staticmethod
    def readLASBatch(filename: str, batch_size: int, process_func, handle=None) -> int:
        """lasio.cpp:211-300: hands the file's points to process_func in batches of batch_size; returns the total."""
        c = PointCloud(handle=handle)
        h = _h(handle)
        hdr = _lib.IcpLasHeader()
        n = C.c_int64()
        if h.lib.icp_las_read(h.h, filename.encode(), 0, VARIANT_ENGINE, C.byref(hdr), None, 0, C.byref(n)) != 0:
            return 0
        pts = np.empty((n.value, 3))
        if h.lib.icp_las_read(h.h, filename.encode(), 0, VARIANT_ENGINE, C.byref(hdr), _ptr(pts), n.value, C.byref(n)) != 0:
            return 0
        total = 0
        for s in range(0, len(pts), int(batch_size)):
            b = pts[s:s + int(batch_size)]
            process_func(b)
            total += len(b)
        return total


def readLASFile(filename: str, cloud: PointCloud, handle=None) -> bool:  # icp_registration.cpp:248-378
    h = _h(handle or cloud._handle)
    hdr = _lib.IcpLasHeader()
    n = C.c_int64()
    if h.lib.icp_las_read(h.h, filename.encode(), 0, VARIANT_CLI, C.byref(hdr), None, 0, C.byref(n)) != 0:
        return False
    pts = np.empty((n.value, 3))
    if h.lib.icp_las_read(h.h, filename.encode(), 0, VARIANT_CLI, C.byref(hdr), _ptr(pts), n.value, C.byref(n)) != 0:
        return False
    cloud.points = pts
    cloud.x_scale, cloud.y_scale, cloud.z_scale = (float(v) for v in hdr.scale)
    cloud.x_offset, cloud.y_offset, cloud.z_offset = (float(v) for v in hdr.offset)
    return n.value > 0


def saveResultAsLAS(cloud: PointCloud, filename: str, handle=None) -> bool:  # icp_registration.cpp:698-815
    h = _h(handle or cloud._handle)
    scale = _v3([cloud.x_scale, cloud.y_scale, cloud.z_scale])
    offset = _v3([cloud.x_offset, cloud.y_offset, cloud.z_offset])
    st = h.lib.icp_las_write(h.h, filename.encode(), _ptr(_c3(cloud.points)), len(cloud.points), VARIANT_CLI, _ptr(scale), _ptr(offset))
    return st == 0


def saveTransformation(R, t, filename: str, iteration_transforms=None) -> bool:  # icp_registration.cpp:625-695
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
    t = _v3(t)
    its = None
    if iteration_transforms is not None and len(iteration_transforms):
        its = np.ascontiguousarray(iteration_transforms, dtype=np.float64).reshape(-1, 16)
    st = _lib.load().icp_save_transformation(filename.encode(), _ptr(R), _ptr(t), _ptr(its), 0 if its is None else len(its))
    return st == 0
