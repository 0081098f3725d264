"""Host-side mirror of the reference's Python API (reference:
hierslam-diff-gaussian-rasterization-w-depth/diff_gaussian_rasterization/__init__.py).

Same public names, argument meaning, return tuples and error behaviour:
``GaussianRasterizationSettings`` (:161-173), ``GaussianRasterizer`` (:175-224),
``GaussianRasterizer_semantic`` (:377-430), ``rasterize_gaussians`` (:20-42),
``rasterize_gaussians_semantic`` (:229-252) and the two autograd functions (:44-159, :254-374).

Differences that stay API-compatible:
  * unmaterialised upstream gradients: ``ctx.set_materialize_grads(False)`` — outputs the loss never touched
    (median depth, silhouette, and during tracking the whole [S,H,W] semantic map) arrive as ``None`` and their
    planes are neither allocated nor read;
  * the work runs on the tensors' device and on torch's *current* stream (the reference uses the legacy default
    stream and has no device guard);
  * S (semantic channels) is a runtime value taken from ``semantics_precomp.shape[1]`` (compile-time constant
    NUM_SEMANTIC in the reference, cuda_rasterizer/config.h:18).
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn

from . import _C


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _call(fn, args, debug: bool, dump: str, what: str):
    """debug=True: snapshot the inputs and dump them if the native call throws (reference :294-301,349-357)."""
    if debug:
        cpu_args = cpu_deep_copy_tuple(args)
        try:
            return fn(*args)
        except Exception as ex:
            torch.save(cpu_args, dump)
            print(f"\nAn error occured in {what}. Please forward {dump} for debugging.")
            raise ex
    return fn(*args)


# ------------------------------------------------------------------------------------------------------
# non-semantic
# ------------------------------------------------------------------------------------------------------
def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        rs = raster_settings
        args = (rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh,
                rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        (num_rendered, color, depth, median_depth, final_opacity, mask, radii, geomBuffer, binningBuffer,
         imgBuffer) = _call(_C.rasterize_gaussians, args, rs.debug, "snapshot_fw.dump", "forward")
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(radii)
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        return color, radii, depth, median_depth, final_opacity, mask

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth, grad_median_depth, grad_final_opacity, grad_mask):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, grad_depth, grad_median_depth,
                grad_final_opacity, sh, rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered, binningBuffer,
                imgBuffer, rs.debug, rs.image_height, rs.image_width)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = _call(_C.rasterize_gaussians_backward, args, rs.debug, "snapshot_bw.dump", "backward")
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales,
                grad_rotations, grad_cov3Ds_precomp, None)


def _check_args(shs, colors_precomp, scales, rotations, cov3D_precomp):
    # same checks and messages as the reference (:196-200, :398-402)
    if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        _check_args(shs, colors_precomp, scales, rotations, cov3D_precomp)
        e = torch.Tensor([])
        return rasterize_gaussians(means3D, means2D, e if shs is None else shs,
                                   e if colors_precomp is None else colors_precomp, opacities,
                                   e if scales is None else scales, e if rotations is None else rotations,
                                   e if cov3D_precomp is None else cov3D_precomp, self.raster_settings)


# ------------------------------------------------------------------------------------------------------
# semantic
# ------------------------------------------------------------------------------------------------------
def rasterize_gaussians_semantic(means3D, means2D, sh, colors_precomp, semantics_precomp, opacities, scales,
                                 rotations, cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians_semantic.apply(means3D, means2D, sh, colors_precomp, semantics_precomp, opacities,
                                              scales, rotations, cov3Ds_precomp, raster_settings)


class _RasterizeGaussians_semantic(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, semantics_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        rs = raster_settings
        args = (rs.bg, means3D, colors_precomp, semantics_precomp, opacities, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
                rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        (num_rendered, color, semantic_map, depth, median_depth, final_opacity, radii, geomBuffer, binningBuffer,
         imgBuffer) = _call(_C.rasterize_gaussians_semantic, args, rs.debug, "snapshot_fw.dump", "forward")
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(radii)
        ctx.save_for_backward(colors_precomp, semantics_precomp, means3D, scales, rotations, cov3Ds_precomp, radii,
                              sh, geomBuffer, binningBuffer, imgBuffer)
        return color, radii, semantic_map, depth, median_depth, final_opacity

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_out_semantic, grad_depth, grad_median_depth,
                 grad_final_opacity):
        rs = ctx.raster_settings
        (colors_precomp, semantics_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
         binningBuffer, imgBuffer) = ctx.saved_tensors
        args = (rs.bg, means3D, radii, colors_precomp, semantics_precomp, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color,
                grad_out_semantic, grad_depth, grad_median_depth, grad_final_opacity, sh, rs.sh_degree, rs.campos,
                geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, rs.debug, rs.image_height, rs.image_width)
        (grad_means2D, grad_colors_precomp, grad_semantics_precomp, grad_opacities, grad_means3D,
         grad_cov3Ds_precomp, grad_sh, grad_scales, grad_rotations) = _call(
            _C.rasterize_gaussians_backward_semantic, args, rs.debug, "snapshot_bw.dump", "backward")
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_semantics_precomp, grad_opacities,
                grad_scales, grad_rotations, grad_cov3Ds_precomp, None)


class GaussianRasterizer_semantic(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, semantics_precomp=None):
        _check_args(shs, colors_precomp, scales, rotations, cov3D_precomp)
        e = torch.Tensor([])
        return rasterize_gaussians_semantic(means3D, means2D, e if shs is None else shs,
                                            e if colors_precomp is None else colors_precomp,
                                            e if semantics_precomp is None else semantics_precomp, opacities,
                                            e if scales is None else scales, e if rotations is None else rotations,
                                            e if cov3D_precomp is None else cov3D_precomp, self.raster_settings)


# ------------------------------------------------------------------------------------------------------
# forward-only depth + silhouette
# ------------------------------------------------------------------------------------------------------
@torch.no_grad()
def render_depth_silhouette(raster_settings, means3D, opacities, scales, rotations):
    """Depth and silhouette of the current map without gradients, colours or semantic channels: what densification
    (scripts/hierslam.py:1307-1352 add_new_gaussians_semantic_newrender: only `rendered_depth` and
    `rendered_final_opcity` of the full semantic render are used) and the evaluation masks need.  Runs the S = 0
    instantiation of the pipeline (4 blended channels instead of 4 + S, no [S,H,W] output) and keeps no autograd state.
    Returns (depth [1,H,W], silhouette [1,H,W], median_depth [1,H,W], radii [P]); bit-identical to the corresponding
    outputs of GaussianRasterizer_semantic on the same inputs."""
    rs = raster_settings
    e = torch.empty(0)
    m = means3D.detach()
    (_, _color, depth, median_depth, final_opacity, _mask, radii, _g, _b, _i) = _C.rasterize_gaussians(
        rs.bg, m, m, opacities.detach(), scales.detach(), rotations.detach(), rs.scale_modifier, e, rs.viewmatrix,
        rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, e, rs.sh_degree, rs.campos, rs.prefiltered,
        rs.debug)
    return depth, final_opacity, median_depth, radii
