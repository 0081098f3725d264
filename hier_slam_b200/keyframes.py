"""Keyframe selection by re-projection (reference utils/keyframe_selection.py:40-96).

The reference back-projects 1600 sampled depth pixels of the current frame and then, PER KEYFRAME, runs ~10 torch kernels
and one host sync to count how many points land inside the keyframe's image.  `keyframe_overlap_counts` counts all
keyframes with one kernel (hs_keyframe_overlap); `keyframe_selection_overlap` mirrors the reference function (same
sampling, sorting and random choice among the overlapping keyframes) around it.  CUDA only."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib


def keyframe_overlap_counts(pts_world: torch.Tensor, est_w2c: torch.Tensor, intrinsics: torch.Tensor, width: int,
                            height: int, edge: int = 20) -> torch.Tensor:
    """pts_world [N,3], est_w2c [K,4,4], intrinsics [3,3] -> int32[K]: points inside each keyframe's image."""
    if not pts_world.is_cuda:
        raise RuntimeError("keyframe_overlap_counts is CUDA-only (no CPU fallback)")
    lib = _lib.load()
    dev = pts_world.device
    pts = pts_world.detach().to(torch.float32).contiguous()
    w2c = est_w2c.detach().to(device=dev, dtype=torch.float32).reshape(-1, 16).contiguous()
    K = w2c.shape[0]
    counts = torch.zeros(K, dtype=torch.int32, device=dev)
    k = intrinsics.detach().to("cpu", torch.float32)
    with torch.cuda.device(dev):
        _lib.check(lib.hs_keyframe_overlap(ctypes.c_void_p(pts.data_ptr()) if pts.numel() else None, int(pts.shape[0]),
                                           ctypes.c_void_p(w2c.data_ptr()) if K else None, K, float(k[0, 0]), float(k[1, 1]),
                                           float(k[0, 2]), float(k[1, 2]), int(width), int(height), int(edge),
                                           ctypes.c_void_p(counts.data_ptr()) if K else None,
                                           ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "hs_keyframe_overlap")
    return counts


def backproject_samples(depth: torch.Tensor, intrinsics: torch.Tensor, w2c: torch.Tensor,
                        sampled_indices: torch.Tensor) -> torch.Tensor:
    """World points of the sampled pixels (reference get_pointcloud, utils/keyframe_selection.py:10-37).  The reference's
    "remove points at the camera origin" step is reproduced as it behaves: after |round(p, 4)| every point that coincides
    with ANOTHER row -- the appended origin, but also a pixel that torch.randint drew twice -- is dropped, all copies."""
    cx, cy, fx, fy = intrinsics[0][2], intrinsics[1][2], intrinsics[0][0], intrinsics[1][1]
    rows, cols = sampled_indices[:, 0], sampled_indices[:, 1]
    z = depth[0, rows, cols]
    cam = torch.stack(((cols - cx) / fx * z, (rows - cy) / fy * z, z), dim=-1)
    pts = (torch.inverse(w2c) @ torch.cat([cam, torch.ones_like(cam[:, :1])], dim=1).T).T[:, :3]
    keyed = torch.cat([torch.abs(torch.round(pts, decimals=4)), torch.zeros((1, 3), dtype=pts.dtype, device=pts.device)], dim=0)
    _, inverse, counts = keyed.unique(dim=0, return_inverse=True, return_counts=True)
    duplicated = counts[inverse][:pts.shape[0]] > 1
    return pts[~duplicated]


def keyframe_selection_overlap(gt_depth, w2c, intrinsics, keyframe_list, k, pixels=1600):
    """Same contract as the reference's keyframe_selection_overlap: ids of up to k random keyframes among those in which
    at least one sampled point of the current frame is visible.  One kernel + one device->host copy for all keyframes."""
    width, height = gt_depth.shape[2], gt_depth.shape[1]
    valid = torch.stack(torch.where(gt_depth[0] > 0), dim=1)
    indices = torch.randint(valid.shape[0], (pixels,))
    pts = backproject_samples(gt_depth, intrinsics, w2c, valid[indices.to(valid.device)])
    if len(keyframe_list) == 0:
        return []
    est = torch.stack([kf['est_w2c'] for kf in keyframe_list])
    counts = keyframe_overlap_counts(pts, est, intrinsics, width, height).cpu().numpy()
    order = sorted(range(len(keyframe_list)), key=lambda i: counts[i], reverse=True)   # stable, like the reference's sort
    selected = [i for i in order if counts[i] > 0]
    return list(np.random.permutation(np.array(selected))[:k])
