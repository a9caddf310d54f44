"""CPU tests of the drop-in boundary: libicp_b200.so loads, exports every symbol include/icp_b200.h declares (and
nothing is declared that the binding does not know), refuses to work without a CUDA device (no CPU fallback), and the
C++ adapters compile and link against it.  No compute calls are made here."""
import ctypes as C
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "icp_b200.h")


def _lib():
    from iterativeclosestpoint_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib


def declared_symbols():
    src = open(HEADER, encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:int|void|int64_t|const char\*)\s+(icp_[a-z_A-Z0-9]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_and_binding_agree():
    lib = _lib()
    decl = declared_symbols()
    assert len(decl) >= 25
    assert sorted(lib.EXPORTED) == decl, (set(decl) ^ set(lib.EXPORTED))


def test_library_exports_every_declared_symbol():
    lib = _lib()
    L = lib.load()
    for name in declared_symbols():
        assert hasattr(L, name), f"libicp_b200.so does not export {name}"
    assert L.icp_abi_version() == 2


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the header structs: sizes a C compiler gives the header's definitions."""
    lib = _lib()
    cc = shutil.which("gcc") or "/usr/bin/gcc"
    prog = ('#include <stdio.h>\n#include "icp_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu ", sizeof(icp_params),'
            'sizeof(icp_iteration), sizeof(icp_stats), sizeof(icp_result), sizeof(icp_octree_info));'
            'printf("%zu %zu\\n", sizeof(icp_las_header), sizeof(icp_las_points));return 0;}\n')
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "s")
        subprocess.check_call([cc, "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(lib.IcpParams), C.sizeof(lib.IcpIteration), C.sizeof(lib.IcpStats), C.sizeof(lib.IcpResult),
                     C.sizeof(lib.IcpOctreeInfo), C.sizeof(lib.IcpLasHeader), C.sizeof(lib.IcpLasPoints)]


def test_host_only_entry_points_work_without_a_device(tmp_path):
    """icp_las_parse_header and icp_save_transformation are host-side framing: no handle, no GPU."""
    import numpy as np
    import io_cases
    from iterativeclosestpoint_b200 import cloudio
    from oracle.binding import OracleIO
    img = io_cases.foreign_las_image()
    h = cloudio.las_parse_header(img[:227])
    assert (h.offset_to_data, h.n_points, h.record_length) == (375, 700, 34)
    assert list(h.scale) == [0.01, 0.01, 0.001] and list(h.offset) == [500_000.0, 4_100_000.0, -12.5]
    bad = img.copy(); bad[0] = ord("X")
    with pytest.raises(_lib().IcpError):
        cloudio.las_parse_header(bad[:227])
    T = io_cases.transform_case(1)
    p = str(tmp_path / "t.txt")
    assert cloudio.saveTransformation(T[:3, :3], T[:3, 3], p, [T])
    assert open(p, "rb").read() == OracleIO().transformation_text(T[:3, :3], T[:3, 3], [T])
    g = np.load(os.path.join(ROOT, "tests", "golden", "io_transformation_text.npz"))
    assert cloudio.saveTransformation(io_cases.transform_case(0)[:3, :3], io_cases.transform_case(0)[:3, 3], p, None)
    assert open(p, "rb").read() == g["text_0"].tobytes()


def test_default_params_are_the_reference_defaults():
    lib = _lib()
    L = lib.load()
    p = lib.IcpParams()
    L.icp_default_params(C.byref(p))
    # ICPParameters defaults, core/icpengine.h:13-19
    assert (p.max_iterations, p.tolerance, p.sigma_multiplier, p.octree_max_points, p.octree_max_depth, p.variant) == (
        50, 1e-6, 3.0, 10, 20, 0)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _lib()
    L = lib.load()
    h = C.c_void_p()
    assert L.icp_create(C.byref(h), 0) == lib.ICP_CUDA_ERROR and not h.value
    from iterativeclosestpoint_b200.engine import Handle, IcpError
    with pytest.raises(IcpError):
        Handle(0)


def test_product_path_never_touches_the_oracle():
    """Nothing under the package may import, link or execute oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "iterativeclosestpoint_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f), encoding="utf-8", errors="replace").read()
                for bad in ("import oracle", "from oracle", "oracle.binding", "liboracle", "icp_oracle", "libref_"):
                    assert bad not in text, f"{f} references {bad}"
    out = subprocess.check_output(["ldd", os.path.join(pkg, "libicp_b200.so")], text=True)
    assert "oracle" not in out and "libref" not in out


def test_cpp_adapters_compile_and_link():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    lib = _lib()
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "adapter_demo")
        subprocess.check_call([cxx, "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tests", "cpp", "adapter_demo.cpp"), "-o", exe,
                               "-L", os.path.dirname(lib.LIB_PATH), "-licp_b200", "-Wl,-rpath," + os.path.dirname(lib.LIB_PATH)])
        assert os.path.exists(exe)
