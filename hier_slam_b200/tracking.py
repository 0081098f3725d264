"""Fused pose step for tracking (north_star: "pre-reduction of ... camera-pose gradients").

Hier-SLAM tracks the camera by moving the Gaussians into the camera frame with torch and letting autograd carry the
rasterizer's dL/dmeans3D [P,3] back to the pose (utils/slam_helpers.py:278-330, scripts/hierslam.py:1837-1852):

    rel_w2c = [R(q) | t];  means_cam = (rel_w2c @ [means_world, 1].T).T[:, :3];  rasterize(means_cam, ...)

which costs a cat, a [4,4]x[4,P] matmul and, in the backward, two more matmuls over [P,3] per iteration.
`PoseRasterizer_semantic` takes the 4x4 pose and the WORLD means instead: the transform is one fused addmm, and the
pose gradient dL/dW[:3,:] = sum_i dL/dmeans_cam_i (x) [means_world_i, 1] is reduced inside the per-Gaussian backward
kernel (warp shuffle -> shared memory -> 12 atomics per block, `geom_backward_kernel`), so no [P,*] autograd node
remains between the pose and the rasterizer.  Outputs and every other gradient are those of
`GaussianRasterizer_semantic` on the transformed means (same kernels, same state).

This is an extension next to the reference-compatible API, not a replacement for it.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _C


class _RasterizePose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w2c, means_world, means2D, colors_precomp, semantics_precomp, opacities, scales, rotations,
                raster_settings):
        rs = raster_settings
        means_cam = torch.addmm(w2c[:3, 3], means_world, w2c[:3, :3].t())      # [P,3], contiguous
        e = torch.Tensor([])
        semantic = semantics_precomp is not None and semantics_precomp.numel() > 0
        if semantic:
            (n, color, sem, depth, median, opacity, radii, gb, bb, ib) = _C.rasterize_gaussians_semantic(
                rs.bg, means_cam, colors_precomp, semantics_precomp, opacities, scales, rotations, rs.scale_modifier, e,
                rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, e, rs.sh_degree,
                rs.campos, rs.prefiltered, rs.debug)
        else:
            (n, color, depth, median, opacity, _mask, radii, gb, bb, ib) = _C.rasterize_gaussians(
                rs.bg, means_cam, colors_precomp, opacities, scales, rotations, rs.scale_modifier, e, rs.viewmatrix,
                rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, e, rs.sh_degree, rs.campos,
                rs.prefiltered, rs.debug)
            sem = torch.zeros(0, rs.image_height, rs.image_width, device=color.device)
        ctx.rs, ctx.num_rendered, ctx.semantic = rs, n, semantic
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(radii)
        ctx.save_for_backward(w2c, means_world, means_cam, colors_precomp, semantics_precomp if semantic else e, scales,
                              rotations, radii, gb, bb, ib)
        return color, radii, sem, depth, median, opacity

    @staticmethod
    def backward(ctx, g_color, g_radii, g_sem, g_depth, g_median, g_opacity):
        rs = ctx.rs
        (w2c, means_world, means_cam, colors, semantics, scales, rotations, radii, gb, bb, ib) = ctx.saved_tensors
        e = torch.Tensor([])
        if ctx.semantic:
            (d_means2D, d_colors, d_sem, d_opac, d_means_cam, _d_cov, _d_sh, d_scales, d_rots, d_pose) = \
                _C.rasterize_gaussians_backward_semantic(
                    rs.bg, means_cam, radii, colors, semantics, scales, rotations, rs.scale_modifier, e, rs.viewmatrix,
                    rs.projmatrix, rs.tanfovx, rs.tanfovy, g_color, g_sem, g_depth, g_median, g_opacity, e, rs.sh_degree,
                    rs.campos, gb, ctx.num_rendered, bb, ib, rs.debug, rs.image_height, rs.image_width,
                    pose_points=means_world)
        else:
            d_sem = None
            (d_means2D, d_colors, d_opac, d_means_cam, _d_cov, _d_sh, d_scales, d_rots, d_pose) = \
                _C.rasterize_gaussians_backward(
                    rs.bg, means_cam, radii, colors, scales, rotations, rs.scale_modifier, e, rs.viewmatrix,
                    rs.projmatrix, rs.tanfovx, rs.tanfovy, g_color, g_depth, g_median, g_opacity, e, rs.sh_degree,
                    rs.campos, gb, ctx.num_rendered, bb, ib, rs.debug, rs.image_height, rs.image_width,
                    pose_points=means_world)
        g_w2c = torch.zeros_like(w2c)
        g_w2c[:3, :] = d_pose
        g_world = d_means_cam @ w2c[:3, :3] if ctx.needs_input_grad[1] else None
        return (g_w2c, g_world, d_means2D, d_colors, d_sem if ctx.semantic else None, d_opac, d_scales, d_rots, None)


class PoseRasterizer_semantic(nn.Module):
    """`GaussianRasterizer_semantic` with the camera pose as an input.

    forward(w2c [4,4], means3D_world [P,3], means2D [P,3], opacities, colors_precomp, scales, rotations,
            semantics_precomp=None) -> (color, radii, semantic, depth, median_depth, final_opacity)

    `raster_settings.viewmatrix / projmatrix` describe the reference frame the pose is relative to (identity / the
    first frame in Hier-SLAM, utils/recon_helpers.py:4-28); `w2c` plays the role of `rel_w2c` of
    `transform_to_frame`.  Precomputed colours and scales + rotations only (what tracking uses)."""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, w2c, means3D, means2D, opacities, colors_precomp, scales, rotations, semantics_precomp=None):
        if w2c.shape != (4, 4):
            raise RuntimeError("w2c must be a [4,4] world-to-camera matrix")
        if means3D.dim() != 2 or means3D.size(1) != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        return _RasterizePose.apply(w2c, means3D.contiguous(), means2D, colors_precomp, semantics_precomp, opacities,
                                    scales, rotations, self.raster_settings)
