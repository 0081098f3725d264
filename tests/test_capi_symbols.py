"""The C-ABI library builds, loads without a GPU and exports every symbol include/hs_raster.h declares."""
import ctypes
import os
import re

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
    assert lib.hs_abi_version() == 3
    for S, ok in ((0, 1), (16, 1), (26, 1), (74, 1), (102, 1), (5, 0), (550, 0)):
        assert lib.hs_supports_semantic_channels(S) == ok


def test_image_state_size_is_host_computable():
    lib = _lib.load()
    n = lib.hs_image_state_bytes(680, 1200)
    assert n >= 680 * 1200 * 8 + 3225 * 8
