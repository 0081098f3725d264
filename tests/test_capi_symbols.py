"""The C-ABI library builds, loads without a GPU and exports every symbol include/hs_raster.h declares."""
import ctypes
import os
import re

import pytest

from hier_slam_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hs_raster.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hs_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_all_declared_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/hs_raster.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms


def test_abi_version_and_supported_channels():
    lib = _lib.load()
    assert lib.hs_abi_version() == 5
    for S, ok in ((0, 1), (16, 1), (26, 1), (74, 1), (102, 1), (5, 0), (550, 0)):
        assert lib.hs_supports_semantic_channels(S) == ok


def test_image_state_size_is_host_computable():
    lib = _lib.load()
    n = lib.hs_image_state_bytes(680, 1200)
    assert n >= 680 * 1200 * 8 + 3225 * 8


def test_header_is_plain_c(tmp_path):
    """include/hs_raster.h is the drop-in boundary of a C ABI: it must compile as C99 (no C++-only syntax), and a C
    translation unit that uses its types, macros and a function must link against libhsraster.so."""
    import shutil
    import subprocess
    from hier_slam_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi_check.c"
    src.write_text('#include "hs_raster.h"\n#include <stdio.h>\n'
                   'int main(void) { hs_camera c; (void)c;\n'
                   '  printf("%d %d %d %d\\n", hs_abi_version(), HS_RASTER_ABI_VERSION, HS_ASYNC_BINNING, HS_POSE_STATE_FLOATS);\n'
                   '  return hs_abi_version() == HS_RASTER_ABI_VERSION ? 0 : 1; }\n')
    exe = tmp_path / "abi_check"
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                        str(src), "-o", str(exe), "-L", lib_dir, "-l:" + os.path.basename(_lib.LIB_PATH),
                        "-Wl,-rpath," + lib_dir, "-Wl,--allow-shlib-undefined"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr     # hs_abi_version() needs no GPU
