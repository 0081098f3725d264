"""TEST INFRASTRUCTURE — torch-CPU restatement of the callers either side of the rasterizer.

Oracle for the widened rows of the hot path: the tracking / mapping losses, the optimiser step, Gaussian pruning, the pose
chain and keyframe selection of Hier-SLAM.  Like ``raster_oracle`` it is NOT part of the product: only ``tests/`` may
import it, and ``hier_slam_b200`` never falls back to it.

Each function cites the reference code it restates (paths relative to /root/reference).  The reference implements these
steps with ordinary torch operators, so the restatements are pinned differently from the CUDA rasterizer's:
``tests/test_slam_oracle.py`` checks every function against an INDEPENDENT formulation (a direct numpy 2-D convolution for
SSIM, a hand-written log-softmax gather for the cross-entropies, torch.optim.Adam itself for the optimiser step, scipy's
quaternion conversion for the pose matrix, numpy loops for the re-projection counts).

Plain float32 torch on whatever device the inputs live on (the GPU parity tests call the same functions on CUDA tensors as
the float32 reference of the kernels)."""
from __future__ import annotations

from math import exp
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


# ---- pose -----------------------------------------------------------------------------------------------------------
def pose_matrix(cam_unnorm_rot: torch.Tensor, cam_tran: torch.Tensor) -> torch.Tensor:
    """rel_w2c of transform_to_frame (utils/slam_helpers.py:278-330): F.normalize of the unnormalised quaternion (r, x, y, z),
    build_rotation (utils/slam_external.py:25-42), translation in the last column."""
    q = F.normalize(cam_unnorm_rot, dim=0)
    r, x, y, z = q[0], q[1], q[2], q[3]
    m = torch.eye(4, dtype=q.dtype, device=q.device)
    m[0, 0], m[0, 1], m[0, 2] = 1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)
    m[1, 0], m[1, 1], m[1, 2] = 2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)
    m[2, 0], m[2, 1], m[2, 2] = 2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)
    m[:3, 3] = cam_tran
    return m


def transform_points(w2c: torch.Tensor, pts_world: torch.Tensor) -> torch.Tensor:
    """(rel_w2c @ [pts, 1].T).T[:, :3] (utils/slam_helpers.py:318-324)."""
    pts4 = torch.cat((pts_world, torch.ones_like(pts_world[:, :1])), dim=1)
    return (w2c @ pts4.T).T[:, :3]


# ---- losses ---------------------------------------------------------------------------------------------------------
def tracking_loss(im, depth, silhouette, gt_im, gt_depth, sil_thres: float = 0.99, use_sil_for_loss: bool = True,
                  depth_weight: float = 1.0, im_weight: float = 0.5) -> torch.Tensor:
    """get_loss_semantic(tracking=True) (scripts/hierslam.py:765-796 with use_l1, no outlier rejection) weighted like
    :1843-1846: masked L1 SUMS of depth and colour; mask = valid depth & finite render [& silhouette > sil_thres]."""
    mask = (gt_depth > 0) & ~torch.isnan(depth)
    if use_sil_for_loss:
        mask = mask & (silhouette > sil_thres)
    mask = mask.detach()
    l_depth = torch.abs(gt_depth - depth)[mask].sum()
    l_im = torch.abs(gt_im - im)[torch.tile(mask, (3, 1, 1))].sum()
    return depth_weight * l_depth + im_weight * l_im


def gaussian_window(window_size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """utils/slam_external.py:55-57."""
    g = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return g / g.sum()


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11) -> torch.Tensor:
    """calc_ssim / _ssim (utils/slam_external.py:60-97): depthwise Gaussian window, zero padding, c1 = 0.01^2, c2 = 0.03^2,
    mean over everything.  img: [C,H,W]."""
    C = img1.shape[0]
    g1 = gaussian_window(window_size).unsqueeze(1)
    win = g1.mm(g1.t()).float()[None, None].expand(C, 1, window_size, window_size).contiguous().to(img1)
    f = lambda x: F.conv2d(x.unsqueeze(0), win, padding=window_size // 2, groups=C).squeeze(0)
    mu1, mu2 = f(img1), f(img2)
    s1, s2, s12 = f(img1 * img1) - mu1 * mu1, f(img2 * img2) - mu2 * mu2, f(img1 * img2) - mu1 * mu2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2))
    return m.mean()


def mapping_colour_loss(im: torch.Tensor, gt_im: torch.Tensor) -> torch.Tensor:
    """scripts/hierslam.py:936: 0.8 * l1_loss_v1(im, gt) + 0.2 * (1 - calc_ssim(im, gt))."""
    return 0.8 * torch.abs(im - gt_im).mean() + 0.2 * (1.0 - ssim(im, gt_im))


def level_cross_entropy(sem: torch.Tensor, labels: torch.Tensor, level_sizes: Sequence[int]) -> torch.Tensor:
    """Inter-level loss (scripts/hierslam.py:955-968 with transfer_tree_rendered_labelmap :91-111): sum over the tree levels
    of CrossEntropyLoss on the level's channel slice.  sem [S,H,W], labels [>= levels, H, W]."""
    ce = torch.nn.CrossEntropyLoss()
    total, beg = 0.0, 0
    for l, n in enumerate(level_sizes):
        total = total + ce(sem[beg:beg + n].permute(1, 2, 0).reshape(-1, n), labels[l].reshape(-1).long())
        beg += n
    return total


def leaf_cross_entropy(sem: torch.Tensor, leaf_labels: torch.Tensor, weight: torch.Tensor,
                       bias: Optional[torch.Tensor]) -> torch.Tensor:
    """Leaf loss (scripts/hierslam.py:975-984): MLP_func = Conv2d(S, classes, 1) (:1756), logits flattened to
    [H*W, classes], CrossEntropyLoss."""
    logits = F.conv2d(sem.unsqueeze(0), weight.reshape(weight.shape[0], -1, 1, 1), bias)
    logits = logits.squeeze(0).view(logits.shape[1], -1).permute(1, 0)
    return torch.nn.CrossEntropyLoss()(logits, leaf_labels.reshape(-1).long())


def tree_semantic_loss(sem, labels, level_sizes, weight, bias, level_weight: float = 1.0, leaf_weight: float = 5.0):
    """losses['sem'] of the tree modes (scripts/hierslam.py:955-984): weight_sem = [1.0, 5.0]; labels [levels + 1, H, W]."""
    return level_weight * level_cross_entropy(sem, labels, level_sizes) + \
        leaf_weight * leaf_cross_entropy(sem, labels[len(level_sizes)], weight, bias)


# ---- optimiser and map maintenance ------------------------------------------------------------------------------------
def adam_step(param, grad, exp_avg, exp_avg_sq, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
    """One torch.optim.Adam update in place (what initialize_optimizer's optimiser does, scripts/hierslam.py:411-417;
    torch/optim/adam.py::_multi_tensor_adam without weight decay / amsgrad); `step` is the 1-based count."""
    b1, b2 = betas
    exp_avg.lerp_(grad, 1 - b1)
    exp_avg_sq.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (exp_avg_sq.sqrt() / (bc2 ** 0.5)).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-(lr / bc1))


def remove_points(tensors: Dict[str, torch.Tensor], to_remove: torch.Tensor) -> Dict[str, torch.Tensor]:
    """remove_points (utils/slam_external.py:142-164): every per-Gaussian tensor indexed with ~to_remove."""
    keep = ~to_remove
    return {k: v[keep] for k, v in tensors.items()}


def prune_mask(logit_opacities, log_scales, threshold: float, scene_radius: float, remove_big: bool) -> torch.Tensor:
    """prune_gaussians' decision (utils/slam_external.py:178-184)."""
    to_remove = (torch.sigmoid(logit_opacities) < threshold).squeeze(-1)
    if remove_big:
        to_remove = torch.logical_or(to_remove, torch.exp(log_scales).max(dim=1).values > 0.1 * scene_radius)
    return to_remove


# ---- keyframe selection ---------------------------------------------------------------------------------------------
def keyframe_overlap_counts(pts_world, est_w2cs, intrinsics, width: int, height: int, edge: int = 20) -> List[int]:
    """The per-keyframe body of keyframe_selection_overlap (utils/keyframe_selection.py:70-88): number of points that
    project inside the keyframe's image (minus the edge margin) with positive depth."""
    out = []
    for est_w2c in est_w2cs:
        t = transform_points(est_w2c, pts_world)
        p2 = torch.matmul(intrinsics, t.transpose(0, 1)).transpose(0, 1)
        pz = p2[:, 2:] + 1e-5
        p2 = (p2 / pz)[:, :2]
        m = (p2[:, 0] < width - edge) * (p2[:, 0] > edge) * (p2[:, 1] < height - edge) * (p2[:, 1] > edge)
        out.append(int((m & (pz[:, 0] > 0)).sum()))
    return out
