"""CPU: the oracle's restatement of the data-format rows (oracle/cloudio_oracle.c) against the golden vectors the
compiled reference produced (tests/golden/io_*.npz, tools/make_golden_io.py) -- bit-exact, bytes and doubles alike --
and, where oracle/_ref exists, against the reference itself on fresh inputs."""
import os

import numpy as np
import pytest

import io_cases
from iterativeclosestpoint_b200 import synth
from oracle import binding
from oracle.binding import OracleIO, VARIANT_CLI, VARIANT_ENGINE

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def same(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


@pytest.fixture(scope="module")
def oio():
    return OracleIO()


def load(name, *inputs):
    g = np.load(os.path.join(GOLD, f"io_{name}.npz"))
    if inputs:
        want = np.frombuffer(bytes.fromhex(synth.digest(*inputs)), dtype=np.uint8)
        assert np.array_equal(g["digest"], want), "input generator drifted from the golden fixture"
    return g


@pytest.mark.parametrize("name", list(io_cases.IO_CLOUDS))
def test_writers_and_readers_match_reference(oio, name):
    make, scale, offset = io_cases.IO_CLOUDS[name]
    xyz = make()
    g = load(name, xyz)
    img = oio.las_file_image(xyz, VARIANT_ENGINE)
    assert same(img, g["engine_image"])                                 # LASIO::writeLAS, every byte
    for mp in io_cases.MAX_POINTS:
        assert same(oio.las_read_image(img, mp), g[f"engine_read_{mp}"])  # LASIO::readLAS(maxPoints)
        mn, mx = oio.bounds(g[f"engine_read_{mp}"])
        assert same(np.r_[mn, mx], g[f"engine_read_bounds_{mp}"])
    assert same(g["engine_batch_points"], g["engine_read_0"]) and list(g["engine_batch_sizes"][:-1]) == [257] * (len(xyz) // 257)
    cimg = oio.las_file_image(xyz, VARIANT_CLI, scale, offset)
    assert same(cimg, g["cli_image"])                                   # saveResultAsLAS
    assert same(oio.las_read_image(cimg), g["cli_read"])                # readLASFile
    assert same(np.asarray(scale, dtype=np.float64), g["cli_read_scale"]) and same(np.asarray(offset, dtype=np.float64), g["cli_read_offset"])
    mn, mx = oio.bounds(xyz)
    assert same(mn, g["bounds_min"]) and same(mx, g["bounds_max"])


@pytest.mark.parametrize("name", list(io_cases.IO_CLOUDS))
def test_downsample_and_apply_match_reference(oio, name):
    xyz = io_cases.IO_CLOUDS[name][0]()
    g = load(name, xyz)
    for t in io_cases.DOWNSAMPLE_TARGETS:
        assert same(oio.downsample(xyz, t), g[f"downsample_{t}"])
    assert g["downsample_null"].all() and len(oio.downsample(xyz, 0)) == 0
    for k in range(2):
        assert same(oio.cloud_apply(io_cases.transform_case(k), xyz), g[f"apply_{k}"])
    for s in io_cases.STRIDES:  # the CLI's sampling is an inline loop of main(): restated, checked against numpy slicing
        assert same(oio.downsample_stride(xyz, s), xyz[::s])


def test_foreign_file_and_failure_exits(oio):
    img = io_cases.foreign_las_image()
    g = load("foreign", img.astype(np.float64))
    assert same(oio.las_read_image(img), g["engine_read"]) and same(oio.las_read_image(img, 100), g["engine_read_100"])
    assert same(oio.las_read_image(img), g["cli_read"])
    bad = img.copy(); bad[:4] = np.frombuffer(b"LASX", dtype=np.uint8)
    assert oio.las_read_image(bad) is None and g["bad_signature_engine_fails"].all() and g["missing_file_fails"].all()
    assert g["bad_signature_cli_reads"].all()  # readLASFile never checks the signature


def test_transformation_text(oio):
    g = np.load(os.path.join(GOLD, "io_transformation_text.npz"))
    for k in range(3):
        T = io_cases.transform_case(k)
        its = np.stack([io_cases.transform_case(j) for j in range(k)]) if k else None
        assert oio.transformation_text(T[:3, :3], T[:3, 3], its) == g[f"text_{k}"].tobytes()


def test_oracle_against_live_reference(oio, tmp_path):
    if not binding.ref_io_available():
        pytest.skip("oracle/_ref not built (needs /root/reference, present only in the build container)")
    ref = binding.RefIO()
    r = np.random.default_rng(2024)
    for trial in range(4):
        n = int(r.integers(1, 4000))
        xyz = r.normal(0, 10.0 ** r.integers(0, 5), (n, 3)) + r.uniform(-1e5, 1e5, 3)
        p = str(tmp_path / f"a{trial}.las")
        assert ref.write_las(p, xyz)
        img = np.fromfile(p, dtype=np.uint8)
        assert same(img, oio.las_file_image(xyz, VARIANT_ENGINE))
        assert same(ref.read_las(p, 0)[0], oio.las_read_image(img))
        scale = 10.0 ** -r.integers(1, 5, 3).astype(np.float64); offset = np.floor(xyz.min(axis=0))
        ref.cli_save_las(p, xyz, scale, offset)
        assert same(np.fromfile(p, dtype=np.uint8), oio.las_file_image(xyz, VARIANT_CLI, scale, offset))
        t = int(r.integers(1, n + 3))
        assert same(ref.downsample(xyz, t), oio.downsample(xyz, t))
