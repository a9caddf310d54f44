"""GPU: the data-format rows of SURVEY.md 8(f) through the C ABI (csrc/cloudio.cu) against the reference's golden vectors
(tests/golden/io_*.npz), against the oracle on fresh inputs, and through size-independent properties at 10 M points.
Bit-exact throughout: every output here is bytes, int32 or a double that the reference computes with two roundings."""
import os

import numpy as np
import pytest

import io_cases
from iterativeclosestpoint_b200 import cloudio, synth
from iterativeclosestpoint_b200._lib import VARIANT_CLI, VARIANT_ENGINE, LAS_HEADER_BYTES, LAS_RECORD_BYTES
from iterativeclosestpoint_b200.engine import ICPParameters
from oracle.binding import OracleIO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def same(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


@pytest.fixture(scope="module")
def oio():
    return OracleIO()


@pytest.mark.parametrize("name", list(io_cases.IO_CLOUDS))
def test_las_files_match_reference(handle, name, tmp_path):
    make, scale, offset = io_cases.IO_CLOUDS[name]
    xyz = make()
    g = np.load(os.path.join(GOLD, f"io_{name}.npz"))
    assert same(cloudio.las_file_image(xyz, VARIANT_ENGINE, handle=handle), g["engine_image"])
    assert same(cloudio.las_file_image(xyz, VARIANT_CLI, scale, offset, handle=handle), g["cli_image"])
    # through the reference-shaped classes and real files
    cloud = cloudio.PointCloud(xyz, handle)
    p = str(tmp_path / "e.las")
    assert cloudio.LASIO.writeLAS(p, cloud)
    assert same(np.fromfile(p, dtype=np.uint8), g["engine_image"])
    for mp in io_cases.MAX_POINTS:
        back = cloudio.PointCloud(handle=handle)
        assert cloudio.LASIO.readLAS(p, back, mp)
        assert same(back.points, g[f"engine_read_{mp}"])
        assert same(np.array([back.minX, back.minY, back.minZ, back.maxX, back.maxY, back.maxZ]), g[f"engine_read_bounds_{mp}"])
    got = []
    assert cloudio.LASIO.readLASBatch(p, 257, lambda b: got.append(b.copy()), handle) == len(xyz)
    assert same(np.concatenate(got), g["engine_batch_points"]) and [len(b) for b in got] == list(g["engine_batch_sizes"])
    cloud.x_scale, cloud.y_scale, cloud.z_scale = scale
    cloud.x_offset, cloud.y_offset, cloud.z_offset = offset
    p2 = str(tmp_path / "c.las")
    assert cloudio.saveResultAsLAS(cloud, p2)
    assert same(np.fromfile(p2, dtype=np.uint8), g["cli_image"])
    back = cloudio.PointCloud(handle=handle)
    assert cloudio.readLASFile(p2, back)
    assert same(back.points, g["cli_read"])
    assert (back.x_scale, back.y_scale, back.z_scale) == tuple(g["cli_read_scale"]) and (back.x_offset, back.y_offset, back.z_offset) == tuple(g["cli_read_offset"])
    mn, mx = cloudio.cloud_bounds(xyz, handle)
    assert same(mn, g["bounds_min"]) and same(mx, g["bounds_max"])


@pytest.mark.parametrize("name", list(io_cases.IO_CLOUDS))
def test_downsample_replay_match_reference(handle, name):
    xyz = io_cases.IO_CLOUDS[name][0]()
    g = np.load(os.path.join(GOLD, f"io_{name}.npz"))
    cloud = cloudio.PointCloud(xyz, handle)
    for t in io_cases.DOWNSAMPLE_TARGETS:
        assert same(cloud.downsample(t).points, g[f"downsample_{t}"])
    assert cloud.downsample(0) is None and cloud.downsample(-5) is None
    for s in io_cases.STRIDES:
        assert same(cloudio.sample_stride(xyz, s, handle), xyz[::s])
    for k in range(2):
        T = io_cases.transform_case(k)
        assert same(cloudio.replay(xyz, T, handle), g[f"apply_{k}"])
        moved = cloudio.PointCloud(xyz.copy(), handle)
        moved.applyTransform(T[:3, :3], T[:3, 3])
        assert same(moved.points, g[f"apply_{k}"])
    assert same(cloudio.replay(xyz, None, handle), xyz)


def test_foreign_file_and_failure_exits(handle, tmp_path):
    img = io_cases.foreign_las_image()
    g = np.load(os.path.join(GOLD, "io_foreign.npz"))
    p = str(tmp_path / "f.las")
    img.tofile(p)
    c = cloudio.PointCloud(handle=handle)
    assert cloudio.LASIO.readLAS(p, c) and same(c.points, g["engine_read"])
    assert cloudio.LASIO.readLAS(p, c, 100) and same(c.points, g["engine_read_100"])
    assert cloudio.readLASFile(p, c) and same(c.points, g["cli_read"])
    bad = img.copy(); bad[:4] = np.frombuffer(b"LASX", dtype=np.uint8)
    pb = str(tmp_path / "bad.las")
    bad.tofile(pb)
    assert not cloudio.LASIO.readLAS(pb, c)                      # lasio.cpp:30-34
    assert cloudio.readLASFile(pb, c) and same(c.points, g["cli_read"])  # the CLI reader does not look at the signature
    assert not cloudio.LASIO.readLAS(str(tmp_path / "missing.las"), c)
    img[:-5].tofile(pb)
    assert not cloudio.LASIO.readLAS(pb, c)                      # truncated point block: reported, not parsed
    assert not cloudio.LASIO.writeLAS(str(tmp_path / "empty.las"), cloudio.PointCloud(handle=handle))  # lasio.cpp:128-131
    assert not cloudio.LASIO.writeLAS(str(tmp_path / "no_such_dir" / "x.las"), cloudio.PointCloud(np.zeros((3, 3)), handle))


def test_transformation_text(tmp_path):
    g = np.load(os.path.join(GOLD, "io_transformation_text.npz"))
    for k in range(3):
        T = io_cases.transform_case(k)
        its = [io_cases.transform_case(j) for j in range(k)]
        p = str(tmp_path / f"t{k}.txt")
        assert cloudio.saveTransformation(T[:3, :3], T[:3, 3], p, its)
        assert open(p, "rb").read() == g[f"text_{k}"].tobytes()


@pytest.mark.parametrize("rl", [20, 26, 28, 34])
def test_decode_encode_vs_oracle_random(handle, oio, rl):
    r = np.random.default_rng(rl)
    n = 200_003
    rec = r.integers(0, 256, n * rl, dtype=np.uint8)
    scale = np.array([0.001, 0.01, 0.00025]); offset = np.array([-431.5, 5.4e6, 12.0])
    xyz = cloudio.las_decode(rec, n, rl, scale, offset, handle)
    assert same(xyz, oio.las_decode(rec, n, rl, scale, offset))
    enc = cloudio.las_encode(xyz * 1.37, scale, offset, handle)
    assert same(enc, oio.las_encode(xyz * 1.37, scale, offset))


def test_full_size_round_trip_properties(handle, oio):
    """10 M points (BASELINE.json config 3's cloud): write -> read -> write is a fixed point after the first pass for the
    CLI writer's fixed scale/offset grid up to the truncation step, records decode to within one scale step, and a
    checksum of the device's records equals the oracle's on the whole cloud."""
    _, tgt = synth.make_pair(10_000_000, 3, "primary")
    scale = np.array([0.001] * 3); offset = np.floor(tgt.min(axis=0))
    rec = cloudio.las_encode(tgt, scale, offset, handle)
    assert rec.size == len(tgt) * LAS_RECORD_BYTES
    want = oio.las_encode(tgt, scale, offset)
    assert int(rec.view(np.uint32).sum(dtype=np.uint64)) == int(want.view(np.uint32).sum(dtype=np.uint64)) and same(rec, want)
    back = cloudio.las_decode(rec, len(tgt), LAS_RECORD_BYTES, scale, offset, handle)
    assert same(back, oio.las_decode(rec, len(tgt), LAS_RECORD_BYTES, scale, offset))
    d = tgt - back
    assert d.min() > -1e-9 and d.max() < 0.001 + 1e-9                 # truncation toward the offset: 0 <= p - decode(encode(p)) < scale
    raw1 = rec.reshape(-1, 5 * 4).view(np.int32)[:, :3]
    raw2 = cloudio.las_encode(back, scale, offset, handle).reshape(-1, 5 * 4).view(np.int32)[:, :3]
    step = raw1 - raw2
    assert step.min() >= 0 and step.max() <= 1                      # re-encoding a decoded point may lose at most one step
    sub = cloudio.downsample(tgt, 1_000_000, handle)
    assert same(sub, tgt[(np.arange(1_000_000) * (len(tgt) / 1_000_000)).astype(np.int32)])
    mn, mx = cloudio.cloud_bounds(tgt, handle)
    assert same(mn, tgt.min(axis=0)) and same(mx, tgt.max(axis=0))


def test_register_from_las_records_equals_register_on_decoded_points(handle, oio):
    """icp_register_las (records decoded on the device) == icp_register on the points LASIO::readLAS returns."""
    src, tgt = synth.make_pair(30_000, 2, "primary")
    simg = oio.las_file_image(src, VARIANT_ENGINE); timg = oio.las_file_image(tgt, VARIANT_ENGINE)
    sh = cloudio.las_parse_header(simg[:LAS_HEADER_BYTES]); th = cloudio.las_parse_header(timg[:LAS_HEADER_BYTES])
    assert (sh.n_points, sh.record_length, sh.offset_to_data) == (len(src), 20, 227)
    s_pts = oio.las_read_image(simg); t_pts = oio.las_read_image(timg)
    handle.set_params(ICPParameters(maxIterations=12))
    work = s_pts.copy()
    want = handle.register(work, t_pts)
    got, moved = cloudio.register_las(simg[LAS_HEADER_BYTES:], sh, timg[LAS_HEADER_BYTES:], th, handle)
    assert got.success == want.success and got.totalIterations == want.totalIterations
    assert same(got.cumulativeT, want.cumulativeT) and got.finalRMSE == want.finalRMSE
    assert same(moved, work)
    ref = oio  # the oracle's ICP on the same decoded clouds: iteration count identical, transform to 1e-9
    from oracle.binding import Oracle
    o = Oracle().icp(s_pts, t_pts, max_iterations=12)
    assert got.totalIterations == o.total_iterations and np.max(np.abs(got.cumulativeT - o.cum_T)) < 1e-9
