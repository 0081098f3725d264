"""Host-side logic of the reference-facing API that needs no GPU: argument validation, error behaviour,
the absence of any CPU fallback, scene determinism."""
import pytest
import torch

import diff_gaussian_rasterization as dgr
from hier_slam_b200.scene import CONFIGS, camera_matrices, keyframe_poses, make_scene, upstream_grads
from oracle import raster_oracle as O


def settings(cfg, device="cpu"):
    view, proj, campos, tfx, tfy = camera_matrices(cfg)
    return dgr.GaussianRasterizationSettings(cfg.height, cfg.width, tfx, tfy, torch.zeros(3), 1.0, view, proj, 0,
                                             campos, False, False)


def test_public_names_match_reference():
    for name in ("GaussianRasterizationSettings", "GaussianRasterizer", "GaussianRasterizer_semantic",
                 "rasterize_gaussians", "rasterize_gaussians_semantic", "_RasterizeGaussians",
                 "_RasterizeGaussians_semantic", "cpu_deep_copy_tuple", "_C"):
        assert hasattr(dgr, name)
    for name in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible",
                 "rasterize_gaussians_semantic", "rasterize_gaussians_backward_semantic"):
        assert hasattr(dgr._C, name)            # reference: ext.cpp:15-23
    assert dgr.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")


def test_argument_validation_messages():
    cfg = CONFIGS["tiny"]
    sc = make_scene(cfg, 0)
    for cls in (dgr.GaussianRasterizer, dgr.GaussianRasterizer_semantic):
        r = cls(settings(cfg))
        with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
            r(means3D=sc["means3D"], means2D=sc["means3D"], opacities=sc["opacities"], scales=sc["scales"],
              rotations=sc["rotations"])
        with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
            r(means3D=sc["means3D"], means2D=sc["means3D"], opacities=sc["opacities"],
              colors_precomp=sc["colors_precomp"])
        with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
            r(means3D=sc["means3D"], means2D=sc["means3D"], opacities=sc["opacities"],
              colors_precomp=sc["colors_precomp"], scales=sc["scales"], rotations=sc["rotations"],
              cov3D_precomp=torch.zeros(sc["means3D"].shape[0], 6))


def test_no_cpu_fallback():
    cfg = CONFIGS["tiny"]
    sc = make_scene(cfg, 0)
    r = dgr.GaussianRasterizer_semantic(settings(cfg))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        r(means3D=sc["means3D"], means2D=sc["means3D"], opacities=sc["opacities"],
          colors_precomp=sc["colors_precomp"], scales=sc["scales"], rotations=sc["rotations"],
          semantics_precomp=sc["semantics_precomp"])
    with pytest.raises(RuntimeError, match="must have dimensions"):
        dgr._C.rasterize_gaussians_semantic(torch.zeros(3), torch.zeros(5, 4), *([torch.zeros(1)] * 5), 1.0,
                                            torch.zeros(0), torch.eye(4), torch.eye(4), 1.0, 1.0, 8, 8,
                                            torch.zeros(0), 0, torch.zeros(3), False, False)


def test_scene_generator_is_seeded():
    cfg = CONFIGS["tiny"]
    a, b, c = make_scene(cfg, 0), make_scene(cfg, 0), make_scene(cfg, 1)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert not torch.equal(a["means3D"], c["means3D"])
    g = upstream_grads(cfg, 1)
    assert g["semantic"].shape == (26, cfg.height, cfg.width)
    assert keyframe_poses(3).shape == (3, 4, 4)
    assert make_scene(CONFIGS["c4"], 0, num_gaussians=10)["semantics_precomp"].shape == (10, 16)


def test_get_higher_msb_matches_documented_key_widths():
    # 44 bits @1200x680 (3225 tiles), 43 @640x480 (1200), 41 @320x240 (300)
    assert 32 + O.get_higher_msb(75 * 43) == 44
    assert 32 + O.get_higher_msb(40 * 30) == 43
    assert 32 + O.get_higher_msb(20 * 15) == 41


def test_padded_channel_counts():
    """Channel counts without their own kernel instantiation map to the next larger one (zero-padded); more than the
    largest instantiation is an error, not a fallback."""
    from hier_slam_b200 import _C
    assert [_C._padded_channels(s) for s in (1, 5, 16, 17, 26, 27, 33, 50, 65, 74, 75, 102)] == [16, 16, 16, 26, 26, 32, 48, 64, 74, 74, 102, 102]
    with pytest.raises(RuntimeError, match="not instantiated"):
        _C._padded_channels(103)


def test_flat_params_cpu_has_no_gradient_sinks():
    """FlatParams on CPU tensors (the gloo tests) must not register CUDA gradient sinks; leaves alias the flat buffers."""
    import torch
    from hier_slam_b200 import _C
    from hier_slam_b200.mapping import FlatParams, keyframes_of_rank
    before = len(_C._grad_sinks)
    p = FlatParams({"a": torch.ones(5, 3), "b": torch.zeros(7)})
    assert len(_C._grad_sinks) == before
    p.leaves["a"].grad.fill_(2.0)
    assert float(p.flat_grad[:15].sum()) == 30.0 and p.flat.numel() % 64 == 0
    assert keyframes_of_rank(8, 1, 4) == [1, 5]


def test_flat_params_empty_layout_and_segments():
    """FlatParams.empty (used by pruning) lays segments out like the constructor: 64-float aligned, leaves alias the flat
    buffers, segment_ends() are the Adam segment boundaries; zero-row tensors are legal."""
    import torch
    from hier_slam_b200.mapping import FlatParams
    a = FlatParams({"x": torch.ones(5, 3), "y": torch.zeros(70, 1), "z": torch.ones(2, 26)})
    b = FlatParams.empty({"x": (5, 3), "y": (70, 1), "z": (2, 26)}, "cpu")
    assert a.offsets == b.offsets == {"x": 0, "y": 64, "z": 192}
    assert a.segment_ends() == b.segment_ends() == [64, 192, 256]
    assert all(e % 4 == 0 for e in b.segment_ends()) and b.flat.numel() == 256
    b.leaves["z"].grad.fill_(1.0)
    assert float(b.flat_grad.sum()) == 52.0 and float(b.flat.sum()) == 0.0
    e = FlatParams.empty({"x": (0, 3), "y": (0, 1)}, "cpu")
    assert e.flat.numel() == 0 and e.leaves["x"].shape == (0, 3)
    b.release()                                   # no CUDA sinks registered: a no-op


def test_flat_adam_is_cuda_only():
    import torch
    from hier_slam_b200.mapping import FlatParams
    from hier_slam_b200.optim import FlatAdam
    with pytest.raises(RuntimeError, match="CUDA-only"):
        FlatAdam(FlatParams({"x": torch.ones(5, 3)}), {"x": 1e-3})


def test_flat_params_appended_matches_torch_cat():
    """FlatParams.appended == torch.cat((param, new_rows)) per parameter (add_new_gaussians / cat_params_to_optimizer),
    in fresh 64-float-aligned buffers; the source set is untouched."""
    import torch
    from hier_slam_b200.mapping import FlatParams
    g = torch.Generator().manual_seed(5)
    a = {"means3D": torch.randn(37, 3, generator=g), "logit_opacities": torch.randn(37, 1, generator=g),
         "semantic": torch.randn(37, 26, generator=g)}
    extra = {"means3D": torch.randn(9, 3, generator=g), "logit_opacities": torch.randn(9, 1, generator=g),
             "semantic": torch.randn(9, 26, generator=g)}
    p = FlatParams({k: v.clone() for k, v in a.items()})
    q = p.appended(extra)
    for k in a:
        assert torch.equal(q.leaves[k].detach(), torch.cat((a[k], extra[k]))) and torch.equal(p.leaves[k].detach(), a[k])
        assert q.leaves[k].grad.shape == q.leaves[k].shape and q.offsets[k] % 64 == 0
    with pytest.raises(RuntimeError, match="same n"):
        p.appended({"means3D": extra["means3D"]})


def test_python_host_has_no_undefined_names():
    """The CUDA-only code paths of the Python host cannot run in the CPU suite; at least every name they load must be
    defined somewhere in its module (catches a helper or constant lost in a refactor)."""
    import ast
    import builtins
    import glob
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hier_slam_b200")
    for path in sorted(glob.glob(os.path.join(root, "*.py"))):
        tree = ast.parse(open(path).read())
        defined = set(dir(builtins)) | {"__file__", "__name__"}
        for n in ast.walk(tree):
            if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
                defined.add(n.name)
            elif isinstance(n, ast.Import):
                defined.update(a.asname or a.name.split(".")[0] for a in n.names)
            elif isinstance(n, ast.ImportFrom):
                defined.update(a.asname or a.name for a in n.names)
            elif isinstance(n, ast.arg):
                defined.add(n.arg)
            elif isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store):
                defined.add(n.id)
            elif isinstance(n, (ast.Global, ast.Nonlocal)):
                defined.update(n.names)
            elif isinstance(n, ast.ExceptHandler) and n.name:
                defined.add(n.name)
        missing = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)
                   and n.id not in defined}
        assert not missing, f"{os.path.basename(path)}: undefined names {sorted(missing)}"


def test_channel_passes_cover_every_channel_once():
    """S beyond the widest instantiation is rendered in passes (hier_slam_b200._C._chunks): the passes tile [0, S) without
    gaps or overlap, each on an instantiated width that holds it."""
    from hier_slam_b200 import _C
    assert _C._chunks(150) == [(0, 74, 74), (74, 74, 74), (148, 2, 16)]
    for S in (103, 148, 149, 550, 1000):
        parts = _C._chunks(S)
        assert parts[0][0] == 0 and sum(n for _, n, _ in parts) == S
        for (c0, n, w), nxt in zip(parts, parts[1:] + [(S, 0, 0)]):
            assert c0 + n == nxt[0] and 0 < n <= w and w in _C._BUILT_S
