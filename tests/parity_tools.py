"""Helpers shared by the GPU parity tests, tools/parity_report.py and bench.py's reference arm:
run either implementation (new CUDA path / reference CUDA build) through the `_C`-level entry points on the
same seeded scene and compare every observable."""
from __future__ import annotations

import os
import sys
from typing import Dict, Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hier_slam_b200.scene import camera_matrices  # noqa: E402


def make_settings(SettingsCls, cfg, device="cuda", w2c=None, debug=False):
    view, proj, campos, tfx, tfy = camera_matrices(cfg, w2c)
    return SettingsCls(image_height=cfg.height, image_width=cfg.width, tanfovx=tfx, tanfovy=tfy,
                       bg=torch.zeros(3, device=device), scale_modifier=1.0, viewmatrix=view.to(device),
                       projmatrix=proj.to(device), sh_degree=0, campos=campos.to(device), prefiltered=False,
                       debug=debug)


def run_forward(C, settings, scene: Dict[str, torch.Tensor], semantic=True):
    """C = a `_C`-compatible module (hier_slam_b200._C or the reference's).  Returns a dict.  This repository's module
    is run with the exact instance count (`exact_num_rendered`: no speculative binning), so that `R` and the state
    buffers can be compared with the reference's; the public-API tests exercise the speculative default."""
    import contextlib
    with (C.exact_num_rendered() if hasattr(C, "exact_num_rendered") else contextlib.nullcontext()):
        return _run_forward(C, settings, scene, semantic)


def _run_forward(C, settings, scene: Dict[str, torch.Tensor], semantic=True):
    e = torch.Tensor([])
    rs = settings
    if semantic:
        args = (rs.bg, scene["means3D"], scene["colors_precomp"], scene["semantics_precomp"], scene["opacities"],
                scene["scales"], scene["rotations"], rs.scale_modifier, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx,
                rs.tanfovy, rs.image_height, rs.image_width, e, 0, rs.campos, False, rs.debug)
        (R, color, sem, depth, median, opacity, radii, gb, bb, ib) = C.rasterize_gaussians_semantic(*args)
        return dict(R=R, color=color, semantic=sem, depth=depth, median_depth=median, final_opacity=opacity,
                    radii=radii, geomBuffer=gb, binningBuffer=bb, imgBuffer=ib)
    args = (rs.bg, scene["means3D"], scene["colors_precomp"], scene["opacities"], scene["scales"], scene["rotations"],
            rs.scale_modifier, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
            rs.image_width, e, 0, rs.campos, False, rs.debug)
    (R, color, depth, median, opacity, mask, radii, gb, bb, ib) = C.rasterize_gaussians(*args)
    return dict(R=R, color=color, depth=depth, median_depth=median, final_opacity=opacity, mask=mask, radii=radii,
                geomBuffer=gb, binningBuffer=bb, imgBuffer=ib)


def run_backward(C, settings, scene, fwd, grads: Dict[str, Optional[torch.Tensor]], semantic=True,
                 materialize=True):
    """grads: color / semantic / depth / median_depth / final_opacity (None allowed when materialize=False,
    which only the new implementation supports; the reference needs real tensors)."""
    e = torch.Tensor([])
    rs = settings
    H, W = rs.image_height, rs.image_width
    dev = scene["means3D"].device

    def g(k, c):
        t = grads.get(k)
        if t is None and materialize:
            return torch.zeros(c, H, W, device=dev)
        return t
    if semantic:
        S = scene["semantics_precomp"].shape[1]
        args = (rs.bg, scene["means3D"], fwd["radii"], scene["colors_precomp"], scene["semantics_precomp"],
                scene["scales"], scene["rotations"], rs.scale_modifier, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx,
                rs.tanfovy, g("color", 3), g("semantic", S), g("depth", 1), g("median_depth", 1),
                g("final_opacity", 1), e, 0, rs.campos, fwd["geomBuffer"], fwd["R"], fwd["binningBuffer"],
                fwd["imgBuffer"], rs.debug)
        if not materialize:
            args = args + (H, W)
        (d_means2D, d_colors, d_sem, d_opac, d_means3D, d_cov3D, d_sh, d_scales, d_rot) = \
            C.rasterize_gaussians_backward_semantic(*args)
        return dict(means2D=d_means2D, colors=d_colors, semantics=d_sem, opacities=d_opac, means3D=d_means3D,
                    cov3D=d_cov3D, scales=d_scales, rotations=d_rot)
    args = (rs.bg, scene["means3D"], fwd["radii"], scene["colors_precomp"], scene["scales"], scene["rotations"],
            rs.scale_modifier, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, g("color", 3), g("depth", 1),
            g("median_depth", 1), g("final_opacity", 1), e, 0, rs.campos, fwd["geomBuffer"], fwd["R"],
            fwd["binningBuffer"], fwd["imgBuffer"], rs.debug)
    if not materialize:
        args = args + (H, W)
    (d_means2D, d_colors, d_opac, d_means3D, d_cov3D, d_sh, d_scales, d_rot) = C.rasterize_gaussians_backward(*args)
    return dict(means2D=d_means2D, colors=d_colors, opacities=d_opac, means3D=d_means3D, cov3D=d_cov3D,
                scales=d_scales, rotations=d_rot)


def bits_equal(a: torch.Tensor, b: torch.Tensor) -> int:
    """number of elements whose BIT PATTERNS differ (float32 compared as int32)."""
    if a.dtype == torch.float32:
        a = a.contiguous().view(torch.int32)
        b = b.contiguous().view(torch.int32)
    return int((a != b).sum().item())


def image_err(a: torch.Tensor, b: torch.Tensor, atol=1e-5, rtol=1e-4):
    """max |a-b|, and the number of pixels violating |a-b| <= atol + rtol |b| (north_star image bar)."""
    d = (a.double() - b.double()).abs()
    viol = int((d > atol + rtol * b.double().abs()).sum().item())
    return float(d.max().item()) if d.numel() else 0.0, viol


def grad_err(a: torch.Tensor, b: torch.Tensor):
    """norm-wise relative error and max element error relative to max|b| (reference atomics are unordered)."""
    a = a.double().reshape(-1)
    b = b.double().reshape(-1)
    nb = float(b.norm().item())
    mb = float(b.abs().max().item()) if b.numel() else 0.0
    d = (a - b)
    return (float(d.norm().item()) / nb if nb > 0 else float(d.norm().item()),
            float(d.abs().max().item()) / mb if mb > 0 else float(d.abs().max().item() if d.numel() else 0.0))
