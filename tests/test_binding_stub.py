"""The reference-side binding INTEGRATION.md section 3 describes is a real file (bindings/torch_ext/hs_torch_ext.cpp: the
five functions of the reference's pybind module, hierslam-diff-gaussian-rasterization-w-depth/ext.cpp:15-23, on top of
include/hs_raster.h).  CPU: it compiles against this torch and this header; when it has been built
(python bindings/torch_ext/build.py) it imports, exports the five names and refuses CPU tensors.  GPU (opt-in,
HS_TEST_NATIVE_BINDING=1 -- the binding was written after the round's GPU budget had ended and has not run on a GPU):
same results as the ctypes host."""
import glob
import importlib.util
import os
import shutil
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIND = os.path.join(ROOT, "bindings", "torch_ext")
NAMES = {"rasterize_gaussians", "rasterize_gaussians_semantic", "rasterize_gaussians_backward",
         "rasterize_gaussians_backward_semantic", "mark_visible"}


def _builder():
    spec = importlib.util.spec_from_file_location("hs_binding_build", os.path.join(BIND, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _built():
    hits = glob.glob(os.path.join(BIND, "_C_native*.so"))
    if not hits:
        return None
    sys.path.insert(0, BIND)
    try:
        import _C_native
        return _C_native
    finally:
        sys.path.remove(BIND)


def test_binding_compiles_against_the_c_abi_header():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    _builder().build(syntax_only=True)        # raises with the compiler output on any mismatch with include/hs_raster.h


def test_built_binding_exports_the_reference_entry_points():
    mod = _built()
    if mod is None:
        pytest.skip("bindings/torch_ext/_C_native*.so not built (python bindings/torch_ext/build.py)")
    assert NAMES <= set(dir(mod))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        mod.mark_visible(torch.zeros(4, 3), torch.eye(4), torch.eye(4))


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("HS_TEST_NATIVE_BINDING") != "1", reason="opt-in: HS_TEST_NATIVE_BINDING=1")
def test_native_binding_matches_the_ctypes_host():
    import parity_tools as pt
    from hier_slam_b200 import _C
    from hier_slam_b200.rasterizer import GaussianRasterizationSettings
    from hier_slam_b200.scene import CONFIGS, make_scene, upstream_grads
    mod = _built()
    if mod is None:
        pytest.skip("binding not built")
    cfg = CONFIGS["small"]
    scene = {k: v.cuda() for k, v in make_scene(cfg, 0).items()}
    settings = pt.make_settings(GaussianRasterizationSettings, cfg)
    ug = {k: v.cuda() for k, v in upstream_grads(cfg, 1).items()}
    fa, fb = pt.run_forward(_C, settings, scene), pt.run_forward(mod, settings, scene)
    assert fa["R"] == fb["R"] and torch.equal(fa["radii"], fb["radii"])
    for k in ("color", "semantic", "depth", "median_depth", "final_opacity"):
        assert torch.equal(fa[k], fb[k]), k
    ga, gb = pt.run_backward(_C, settings, scene, fa, ug), pt.run_backward(mod, settings, scene, fb, ug)
    for k in ga:
        if ga[k] is not None:
            nrm, _ = pt.grad_err(gb[k], ga[k])
            assert nrm < 1e-5, (k, nrm)
