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
