"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference CUDA extension built by oracle/build_ref.sh.

The reference fixes its semantic channel count at compile time (cuda_rasterizer/config.h:18), so there is one
build per S under oracle/_ref/S<S>/.  Each build is imported under its own alias (``hsref_S<S>``) so several
variants and the new implementation (which owns the module name ``diff_gaussian_rasterization`` in this repo)
can live in one process.

Also parses the reference's three opaque byte buffers (layout: cuda_rasterizer/rasterizer_impl.cu:155-194,
rasterizer_impl.h:21-73) so the parity tests can compare keys / lists / ranges / n_contrib bit for bit.
"""
from __future__ import annotations

import glob
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available(S: int) -> bool:
    return bool(glob.glob(os.path.join(REF_ROOT, f"S{S}", "diff_gaussian_rasterization", "_C*.so")))


def load_reference(S: int):
    """Returns the reference's python package (its own __init__.py + its own _C) for NUM_SEMANTIC == S, or None."""
    name = f"hsref_S{S}"
    if name in sys.modules:
        return sys.modules[name]
    if not available(S):
        return None
    pkg_dir = os.path.join(REF_ROOT, f"S{S}", "diff_gaussian_rasterization")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        del sys.modules[name]
        raise
    return mod


def _align(x: int, a: int = 128) -> int:
    return (x + a - 1) & ~(a - 1)


def parse_ref_state(P: int, H: int, W: int, R: int, geomBuffer: torch.Tensor, binningBuffer: torch.Tensor,
                    imgBuffer: torch.Tensor):
    """Typed views into the reference's geom / binning / img buffers.  Offsets are relative to the buffer start,
    which torch allocates 512-B aligned, so the reference's address-based 128-B alignment equals offset alignment.
    Every array used here sits before the CUB temporary storage, whose size is toolkit-dependent."""
    N = H * W

    def take(buf, off, nbytes, dtype):
        off = _align(off)
        return buf[off:off + nbytes].view(dtype), off + nbytes
    out = {}
    o = 0
    out["depths"], o = take(geomBuffer, o, 4 * P, torch.float32)
    _, o = take(geomBuffer, o, 3 * P, torch.uint8)                      # clamped bool[3P]
    out["internal_radii"], o = take(geomBuffer, o, 4 * P, torch.int32)
    m2d, o = take(geomBuffer, o, 8 * P, torch.float32)
    out["means2D"] = m2d.view(P, 2)
    c3, o = take(geomBuffer, o, 24 * P, torch.float32)
    out["cov3D"] = c3.view(P, 6)
    co, o = take(geomBuffer, o, 16 * P, torch.float32)
    out["conic_opacity"] = co.view(P, 4)
    _, o = take(geomBuffer, o, 12 * P, torch.float32)                   # rgb
    out["tiles_touched"], o = take(geomBuffer, o, 4 * P, torch.int32)
    o = 0
    out["final_T"], o = take(imgBuffer, o, 4 * N, torch.float32)
    out["n_contrib"], o = take(imgBuffer, o, 4 * N, torch.int32)
    rg, o = take(imgBuffer, o, 8 * N, torch.int32)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    out["ranges"] = rg.view(N, 2)[:tiles]
    if R > 0:
        o = 0
        out["point_list"], o = take(binningBuffer, o, 4 * R, torch.int32)
        out["point_list_unsorted"], o = take(binningBuffer, o, 4 * R, torch.int32)
        out["keys"], o = take(binningBuffer, o, 8 * R, torch.int64)
        out["keys_unsorted"], o = take(binningBuffer, o, 8 * R, torch.int64)
    return out
