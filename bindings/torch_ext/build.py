"""Builds bindings/torch_ext/hs_torch_ext.cpp -- the pybind11 / torch C++ form of the reference's native module over the
C ABI -- into bindings/torch_ext/_C_native*.so with g++ (no nvcc: the file contains no device code; the kernels live in
hier_slam_b200/libhsraster.so, which it links).

    python bindings/torch_ext/build.py [--syntax-only]

To use it under the reference's own Python layer: `from <this dir> import _C_native as _C` in place of `from . import _C`
(hierslam-diff-gaussian-rasterization-w-depth/diff_gaussian_rasterization/__init__.py:15)."""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "hs_torch_ext.cpp")
NAME = "_C_native"


def command(syntax_only: bool = False):
    import torch
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], os.path.join(ROOT, "include"),
                                os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")]
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"] + [f"-I{i}" for i in inc]
    if syntax_only:
        return cmd + ["-fsyntax-only", SRC], None
    out = os.path.join(HERE, NAME + sysconfig.get_config_var("EXT_SUFFIX"))
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    hlib = os.path.join(ROOT, "hier_slam_b200")
    cmd += ["-shared", SRC, "-o", out, f"-L{tlib}", f"-L{hlib}", "-lhsraster", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10",
            "-lc10_cuda", f"-Wl,-rpath,{tlib}", f"-Wl,-rpath,{hlib}"]
    return cmd, out


def build(syntax_only: bool = False):
    cmd, out = command(syntax_only)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stderr[-4000:])
    return out


if __name__ == "__main__":
    print(build("--syntax-only" in sys.argv) or "syntax ok")
