"""Loss helpers next to the rasterizer (first step): `masked_l1_sum`.

Hier-SLAM's tracking and mapping losses are `torch.abs(gt - x)[mask].sum()` (scripts/hierslam.py:780-796); the boolean
indexing costs a nonzero() with a host sync in the forward and an index_put_ in the backward.  `masked_l1_sum` computes
the same value and the same gradient with one kernel over the image (hs_masked_l1)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class _MaskedL1Sum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask):
        lib = _lib.load()
        if not pred.is_cuda:
            raise RuntimeError("masked_l1_sum is CUDA-only (no CPU fallback)")
        if pred.dtype != torch.float32 or target.dtype != torch.float32 or pred.shape != target.shape:
            raise RuntimeError("pred and target must be float32 tensors of the same shape [C,H,W]")
        p, t = pred.contiguous(), target.contiguous()
        C, HW = (p.shape[0], p[0].numel()) if p.dim() == 3 else (1, p.numel())
        m = None
        if mask is not None:
            m = mask.reshape(-1)
            if m.numel() != HW:
                raise RuntimeError("mask must have one entry per pixel ([H,W] or [1,H,W])")
            m = m.contiguous().view(torch.uint8) if m.dtype == torch.bool else (m != 0).view(torch.uint8)
        loss = torch.zeros((), dtype=torch.float32, device=p.device)
        grad = torch.empty_like(p)
        with torch.cuda.device(p.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
            _lib.check(lib.hs_masked_l1(ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(t.data_ptr()),
                                        ctypes.c_void_p(m.data_ptr()) if m is not None else None, int(C), int(HW),
                                        ctypes.c_void_p(loss.data_ptr()), ctypes.c_void_p(grad.data_ptr()), stream),
                       "hs_masked_l1")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def masked_l1_sum(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
    """sum over channels and masked pixels of |pred - target|  ==  torch.abs(target - pred)[mask_expanded].sum().
    pred / target: [C,H,W] float32 CUDA; mask: [H,W] or [1,H,W] bool (None = all pixels).  Gradient flows to pred."""
    return _MaskedL1Sum.apply(pred, target, mask)


class _HierCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sem, labels, level_sizes, weights):
        lib = _lib.load()
        if not sem.is_cuda or sem.dtype != torch.float32 or sem.dim() != 3:
            raise RuntimeError("sem must be a float32 CUDA tensor [S,H,W] (no CPU fallback)")
        L = len(level_sizes)
        if labels.dim() != 3 or labels.shape[0] < L or labels.shape[1:] != sem.shape[1:]:
            raise RuntimeError("labels must be [levels,H,W] with at least len(level_sizes) levels")
        if sum(level_sizes) > sem.shape[0] or L > 8:
            raise RuntimeError("level_sizes must sum to at most S and have at most 8 levels")
        s = sem.contiguous()
        lab = labels[:L].to(torch.int32).contiguous()
        HW = s[0].numel()
        counts = (lab >= 0).reshape(L, -1).sum(1).clamp_min(1).tolist()     # torch's mean over the non-ignored pixels
        begin = (ctypes.c_int * (L + 1))(*([0] + [sum(level_sizes[:i + 1]) for i in range(L)]))
        scale = (ctypes.c_float * L)(*[float(weights[i]) / counts[i] for i in range(L)])
        loss = torch.zeros((), dtype=torch.float32, device=s.device)
        grad = torch.zeros_like(s) if sum(level_sizes) < s.shape[0] else torch.empty_like(s)
        with torch.cuda.device(s.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(s.device).cuda_stream)
            _lib.check(lib.hs_hier_cross_entropy(ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(lab.data_ptr()), L, begin,
                                                 scale, HW, ctypes.c_void_p(loss.data_ptr()),
                                                 ctypes.c_void_p(grad.data_ptr()), stream), "hs_hier_cross_entropy")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


def hierarchical_cross_entropy(sem: torch.Tensor, labels: torch.Tensor, level_sizes, weights=None) -> torch.Tensor:
    """sum_l weights[l] * CrossEntropyLoss()(sem[begin_l:end_l].permute(1,2,0).view(-1, n_l), labels[l].view(-1))
    -- the inter-level loss of scripts/hierslam.py:955-1000 -- for the planar semantic map [S,H,W] the rasterizer
    returns, in one kernel, with the gradient written in the same planar layout.  labels: [levels,H,W] integer
    (negative = ignored, like torch's ignore_index)."""
    level_sizes = [int(v) for v in level_sizes]
    weights = [1.0] * len(level_sizes) if weights is None else [float(w) for w in weights]
    return _HierCE.apply(sem, labels, level_sizes, weights)
