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


# ------------------------------------------------------------------------------------------------------
# CUDA-graph tracking loop
# ------------------------------------------------------------------------------------------------------
def _pose_matrix(cam_rot: torch.Tensor, cam_tran: torch.Tensor) -> torch.Tensor:
    """rel_w2c of transform_to_frame (utils/slam_helpers.py:278-330): normalised quaternion (r, x, y, z) -> rotation
    (utils/slam_external.py build_rotation), translation in the last column."""
    q = torch.nn.functional.normalize(cam_rot, dim=0)
    r, x, y, z = q[0], q[1], q[2], q[3]
    one, zero = torch.ones_like(r), torch.zeros_like(r)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y), cam_tran[0]]),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x), cam_tran[1]]),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y), cam_tran[2]]),
        torch.stack([zero, zero, zero, one])])


class GraphedTracker:
    """Hier-SLAM's per-frame camera tracking (scripts/hierslam.py:1805-1880; BASELINE.json config 3) with ONE CUDA-graph
    launch per iteration.

    An iteration is: pose -> rel_w2c -> camera-frame means -> render (S = 0 pipeline: colour + depth + silhouette; the
    semantic channels are not rendered because no tracking loss reads them) -> mask = (gt_depth > 0) & ~isnan(depth)
    [& (silhouette > sil_thres)] -> loss = depth_weight * sum|gt_depth - depth|[mask] + im_weight * sum|gt_im - im|[mask]
    (get_loss_semantic(tracking=True), :765-796) -> best-candidate bookkeeping (:1850-1856) -> backward with the pose
    gradient reduced in the per-Gaussian kernel -> Adam on the unnormalised quaternion and the translation
    (configs/replica/hierslam_semantic_run.py:85-95) -> best-candidate bookkeeping (:1851-1858: the pose after the step is
    kept when the loss before it was the smallest so far).

    Nothing in it needs autograd or the host: the steps between the rasterizer calls are three kernels of the library
    (hs_transform_points, hs_tracking_loss, hs_pose_step; in torch they are ~200 tiny launches per iteration) and the
    forward runs in capacity mode (`_C.BinningCapacity`: no `num_rendered` read-back), so an iteration is ~16 launches
    replayed from one graph.  The overflow flag is read ONCE per frame together with the result; a frame that outgrew
    the capacity is repeated after re-capturing with a larger one.  The graph is re-captured when the number of
    Gaussians changes."""

    def __init__(self, raster_settings, lr_rot: float = 0.0004, lr_trans: float = 0.002, sil_thres: float = 0.99,
                 use_sil_for_loss: bool = True, depth_weight: float = 1.0, im_weight: float = 0.5, slack: float = 1.3,
                 extra_instances: int = 65536, betas=(0.9, 0.999), eps: float = 1e-8):
        self.rs = raster_settings
        self.lr_rot, self.lr_trans = float(lr_rot), float(lr_trans)
        self.sil_thres, self.use_sil = float(sil_thres), bool(use_sil_for_loss)
        self.depth_weight, self.im_weight = float(depth_weight), float(im_weight)
        self.slack, self.extra_instances = float(slack), int(extra_instances)
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.graph = None
        self.capacity = None
        self.captures = 0
        self._P = None

    def _pose_step(self, mode: int, d_pose=None, info=None):
        import ctypes
        from . import _lib
        st = self.st
        vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        dev = st["cam_rot"].device
        _lib.check(_lib.load().hs_pose_step(vp(st["cam_rot"]), vp(st["cam_tran"]), vp(d_pose), vp(st["loss"]),
                                            vp(st["state"]), vp(st["w2c"]), vp(info), self.lr_rot, self.lr_trans,
                                            self.betas[0], self.betas[1], self.eps, mode,
                                            ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "hs_pose_step")

    # -- one iteration on the static buffers (runs eagerly for warm-up, and once under capture); returns the image buffer
    @torch.no_grad()
    def _iteration(self):
        import ctypes
        from . import _lib
        lib, st, rs = _lib.load(), self.st, self.cam
        dev = st["means3D"].device
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        H, W, P = rs.image_height, rs.image_width, st["means3D"].shape[0]
        e = torch.empty(0)
        _lib.check(lib.hs_transform_points(vp(st["w2c"]), vp(st["means3D"]), P, vp(st["means_cam"]), stream),
                   "hs_transform_points")
        (n, im, depth, _median, sil, _mask, radii, gb, bb, ib) = _C.rasterize_gaussians(
            rs.bg, st["means_cam"], st["rgb_colors"], st["opacities"], st["scales"], st["rotations"], rs.scale_modifier, e,
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, H, W, e, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        g_im, g_depth = torch.empty_like(im), torch.empty_like(depth)
        _lib.check(lib.hs_tracking_loss(vp(im), vp(depth), vp(sil), vp(st["gt_im"]), vp(st["gt_depth"]), H * W,
                                        self.sil_thres, int(self.use_sil), self.depth_weight, self.im_weight,
                                        vp(st["loss"]), vp(g_im), vp(g_depth), stream), "hs_tracking_loss")
        grads = _C.rasterize_gaussians_backward(
            rs.bg, st["means_cam"], radii, st["rgb_colors"], st["scales"], st["rotations"], rs.scale_modifier, e,
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, g_im, g_depth, None, None, e, rs.sh_degree, rs.campos,
            gb, n, bb, ib, rs.debug, H, W, pose_points=st["means3D"])
        # the forward's binning counts go to the pose step: an iteration that outgrew the capacity is discarded there and
        # latched into the pose state (the flag itself is rewritten by every forward)
        self._pose_step(1, grads[-1], _C.binning_info(ib, H, W))
        return ib

    def _allocate(self, P: int, dev):
        H, W = self.rs.image_height, self.rs.image_width
        f = dict(dtype=torch.float32, device=dev)
        self.st = dict(means3D=torch.zeros(P, 3, **f), means_cam=torch.zeros(P, 3, **f), opacities=torch.zeros(P, 1, **f),
                       rgb_colors=torch.zeros(P, 3, **f), scales=torch.zeros(P, 3, **f), rotations=torch.zeros(P, 4, **f),
                       gt_im=torch.zeros(3, H, W, **f), gt_depth=torch.zeros(1, H, W, **f), cam_rot=torch.zeros(4, **f),
                       cam_tran=torch.zeros(3, **f), w2c=torch.zeros(16, **f), loss=torch.zeros(1, **f),
                       state=torch.zeros(32, **f))
        # Camera tensors whose addresses are baked into the captured graph are OWNED here (contiguous float32 copies that
        # live as long as the tracker); track() refreshes their contents from the caller's settings before every frame,
        # so in-place edits of rs.viewmatrix / rs.projmatrix are honoured and nothing the graph reads can be freed.
        rs = self.rs
        own = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous().clone()
        self.cam = rs._replace(bg=own(rs.bg), viewmatrix=own(rs.viewmatrix), projmatrix=own(rs.projmatrix),
                               campos=own(rs.campos))
        self._P = P
        self.graph = None

    @torch.no_grad()
    def _reset(self, init_rot, init_tran):
        st = self.st
        st["cam_rot"].copy_(init_rot.reshape(4))
        st["cam_tran"].copy_(init_tran.reshape(3))
        st["state"].zero_()                    # the reference builds a fresh optimizer for every frame
        st["state"][15] = 1e20                 # min_loss
        st["state"][16:20].copy_(st["cam_rot"])
        st["state"][20:23].copy_(st["cam_tran"])
        st["loss"].zero_()
        self._pose_step(0)

    def _capture(self, at_least=None):
        """eager synchronous iterations (warm-up + the counts that size the capacity), then the capture"""
        dev = self.st["means3D"].device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                ib = self._iteration()
            info = _C.binning_info(ib, self.rs.image_height, self.rs.image_width)
            cap = _C.BinningCapacity.from_info(info, self.slack, self.extra_instances)
            if at_least is not None:      # what an overflowed frame actually needed, with head-room
                cap = _C.BinningCapacity(max(cap.instances, at_least.instances),
                                         max(cap.longest_tile, at_least.longest_tile))
        torch.cuda.current_stream(dev).wait_stream(side)
        self.capacity = cap
        self.graph = torch.cuda.CUDAGraph()
        with _C.async_binning(cap):
            with torch.cuda.graph(self.graph):
                self._iteration()
        self.captures += 1

    @torch.no_grad()
    def _load(self, means3D, rgb_colors, opacities, scales, rotations, gt_im, gt_depth):
        st = self.st
        for k, v in (("means3D", means3D), ("rgb_colors", rgb_colors), ("opacities", opacities), ("scales", scales),
                     ("rotations", rotations), ("gt_im", gt_im), ("gt_depth", gt_depth)):
            st[k].copy_(v.detach().reshape(st[k].shape))
        for k in ("bg", "viewmatrix", "projmatrix", "campos"):
            getattr(self.cam, k).copy_(getattr(self.rs, k).detach().reshape(getattr(self.cam, k).shape))

    def track(self, means3D, rgb_colors, opacities, scales, rotations, gt_im, gt_depth, init_rot, init_tran,
              num_iters: int = 40, max_retries: int = 3):
        """Optimise the pose of one frame.  means3D [P,3] are WORLD-frame means; opacities / scales [P,3] / rotations are
        the render variables of transformed_params2rendervar (activated, normalised); gt_im [3,H,W], gt_depth [1,H,W].
        Returns dict(rot, tran: the best candidate; loss: its loss; last_rot, last_tran, last_loss; retries)."""
        if not means3D.is_cuda:
            raise RuntimeError("GraphedTracker is CUDA-only (no CPU fallback)")
        P, dev = means3D.shape[0], means3D.device
        with torch.cuda.device(dev):
            if self._P != P or self.st["means3D"].device != dev:
                self._allocate(P, dev)
            self._load(means3D, rgb_colors, opacities, scales, rotations, gt_im, gt_depth)
            retries, at_least = 0, None
            while True:
                self._reset(init_rot, init_tran)
                if self.graph is None:
                    self._capture(at_least)
                    self._reset(init_rot, init_tran)
                for _ in range(num_iters):
                    self.graph.replay()
                st = self.st
                # state[24]: overflow latch of the WHOLE frame (any iteration), written by hs_pose_step
                host = torch.cat((st["state"], st["cam_rot"], st["cam_tran"])).cpu()     # the frame's one sync
                if host[24] == 0 or retries >= max_retries:
                    break
                retries += 1          # some iteration outgrew the binning capacity: re-capture larger and repeat the frame
                # state[25:27]: the largest counts any iteration of the frame needed (an overflowed forward still reports them)
                at_least = _C.BinningCapacity(int(float(host[25]) * 1.3) + 65536, int(float(host[26]) * 1.3) + 256)
                self.graph = None
            if host[24] != 0:
                raise RuntimeError("tracking frame does not fit the binning capacity after re-captures")
        return dict(rot=host[16:20].clone(), tran=host[20:23].clone(), loss=float(host[15]), last_rot=host[32:36].clone(),
                    last_tran=host[36:39].clone(), last_loss=float(host[23]), retries=retries)
