"""Loss kernels next to the rasterizer: `masked_l1_sum`, `hierarchical_cross_entropy`,
`leaf_cross_entropy`, `tree_semantic_loss`, `l1_ssim_loss`.

Hier-SLAM's tracking and mapping losses are `torch.abs(gt - x)[mask].sum()` (scripts/hierslam.py:780-796); the boolean
indexing costs a nonzero() with a host sync in the forward and an index_put_ in the backward.  `masked_l1_sum` computes
the same value and the same gradient with one kernel over the image (hs_masked_l1)."""
from __future__ import annotations

import ctypes
import os

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
    def forward(ctx, sem, labels, level_sizes, weights, level_valid):
        lib = _lib.load()
        if not sem.is_cuda or sem.dtype != torch.float32 or sem.dim() != 3:
            raise RuntimeError("sem must be a float32 CUDA tensor [S,H,W] (no CPU fallback)")
        L = len(level_sizes)
        if labels.dim() != 3 or labels.shape[0] < L or labels.shape[1:] != sem.shape[1:]:
            raise RuntimeError("labels must be [levels,H,W] with at least len(level_sizes) levels")
        if sum(level_sizes) > sem.shape[0] or L > 8:
            raise RuntimeError("level_sizes must sum to at most S and have at most 8 levels")
        s = sem.detach().contiguous()
        lab = labels[:L].to(torch.int32).contiguous()           # a no-op for int32 labels (convert once per keyframe)
        loss = torch.zeros((), dtype=torch.float32, device=s.device)
        grad = torch.zeros_like(s) if sum(level_sizes) < s.shape[0] else torch.empty_like(s)
        with torch.cuda.device(s.device):
            _run_hier(lib, s, lab, level_sizes, weights, loss, grad, level_valid)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None


def hierarchical_cross_entropy(sem: torch.Tensor, labels: torch.Tensor, level_sizes, weights=None,
                               level_valid=None) -> torch.Tensor:
    """sum_l weights[l] * CrossEntropyLoss()(sem[begin_l:end_l].permute(1,2,0).view(-1, n_l), labels[l].view(-1))
    -- the inter-level loss of scripts/hierslam.py:955-1000 -- for the planar semantic map [S,H,W] the rasterizer
    returns, in one kernel, with the gradient written in the same planar layout.  labels: [levels,H,W] integer
    (negative = ignored, like torch's ignore_index; int32 labels are used as they are).  level_valid: the number of
    non-ignored pixels per level if the caller knows it (H*W without ignored labels) -- saves a host sync."""
    level_sizes = [int(v) for v in level_sizes]
    weights = [1.0] * len(level_sizes) if weights is None else [float(w) for w in weights]
    return _HierCE.apply(sem, labels, level_sizes, weights, level_valid)


# HS_LEAF_TF32=1: one TF32 product per contraction in the leaf loss (torch's default convolution precision) instead of the
# fp32-accurate 3xTF32 split
LEAF_TF32 = os.environ.get("HS_LEAF_TF32", "0") == "1"
# Per-pixel pass of the leaf loss: two kernels, both fp32-accurate 3xTF32, chosen by the class count.
#   csrc/leaf_loss_tc.cu  tcgen05 / TMEM / TMA: classes stream through tensor memory in chunks of 64 -- any class count;
#                         74 -> 550 classes at 640x480: 0.37 ms vs 1.94 ms (64 % of the measured TF32 peak, 3 products counted)
#   csrc/leaf_loss.cu     mma.sync one-pass: all logits of a 16-pixel row tile stay in a warp's registers -- only up to 112
#                         classes, where it is the faster one (26 -> 102 at 1200x680: 0.26 ms vs 0.30 ms)
# LEAF_KERNEL / HS_LEAF_KERNEL = "auto" | "tcgen05" | "mma_sync" forces one of them (A/B measurements, tests).
LEAF_KERNEL = os.environ.get("HS_LEAF_KERNEL", "auto")
LEAF_TC_MIN_CLASSES = 113


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _run_hier(lib, s, lab, level_sizes, weights, loss, grad, level_valid=None):
    """hs_hier_cross_entropy on contiguous sem [S,H,W] / int32 labels [L,H,W]; writes grad[: sum(level_sizes)]."""
    L = len(level_sizes)
    if level_valid is None:
        counts = (lab >= 0).reshape(L, -1).sum(1).clamp_min(1).tolist()  # torch's mean over the non-ignored pixels (host sync)
    else:
        counts = [max(int(v), 1) for v in ([level_valid] * L if isinstance(level_valid, int) else level_valid)]
    begin = (ctypes.c_int * (L + 1))(*([0] + [sum(level_sizes[:i + 1]) for i in range(L)]))
    scale = (ctypes.c_float * L)(*[float(weights[i]) / counts[i] for i in range(L)])
    _lib.check(lib.hs_hier_cross_entropy(ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(lab.data_ptr()), L, begin, scale,
                                         s[0].numel(), ctypes.c_void_p(loss.data_ptr()), ctypes.c_void_p(grad.data_ptr()),
                                         _stream(s.device)), "hs_hier_cross_entropy")


def _run_leaf(lib, s, lab, w2, bias, loss_weight, num_valid, loss, grad, accumulate, want_wgrad):
    """hs_leaf_cross_entropy; returns (grad_weight [L,S] or None, grad_bias [L] or None)."""
    L, S = w2.shape
    if S != s.shape[0]:
        raise RuntimeError(f"weight has {S} input channels, the semantic map has {s.shape[0]}")
    if not 1 <= S <= 79:
        raise RuntimeError("leaf_cross_entropy supports 1 <= S <= 79 semantic channels")
    if num_valid is None:
        num_valid = int(((lab >= 0) & (lab < L)).sum())                 # host sync; pass num_valid to avoid it
    HW = s[0].numel()
    lse = torch.empty(HW, dtype=torch.float32, device=s.device)
    gw = torch.zeros(L, S, dtype=torch.float32, device=s.device) if want_wgrad else None
    gb = torch.zeros(L, dtype=torch.float32, device=s.device) if (want_wgrad and bias is not None) else None
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    flags = (1 if accumulate else 0) | (2 if LEAF_TF32 else 0)
    use_tc = LEAF_KERNEL == "tcgen05" or (LEAF_KERNEL == "auto" and L >= LEAF_TC_MIN_CLASSES)
    if not use_tc:
        _lib.check(lib.hs_leaf_cross_entropy(p(s), p(lab), p(w2), p(bias), int(S), int(L), HW,
                                             float(loss_weight) / max(int(num_valid), 1), p(loss), p(lse), p(grad),
                                             flags, p(gw), p(gb), _stream(s.device)), "hs_leaf_cross_entropy")
        return gw, gb
    nbytes = int(lib.hs_leaf_ce_workspace_bytes(int(S), int(L)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=s.device)     # caching allocator: 512-byte aligned
    _lib.check(lib.hs_leaf_cross_entropy_tc(p(s), p(lab), p(w2), p(bias), int(S), int(L), HW,
                                            float(loss_weight) / max(int(num_valid), 1), p(loss), p(lse), p(grad),
                                            flags, p(gw), p(gb), p(ws), nbytes, _stream(s.device)),
               "hs_leaf_cross_entropy_tc")
    return gw, gb


def _leaf_args(sem, weight, bias):
    if not sem.is_cuda or sem.dtype != torch.float32 or sem.dim() != 3:
        raise RuntimeError("sem must be a float32 CUDA tensor [S,H,W] (no CPU fallback)")
    if weight.dim() not in (2, 4) or weight.dtype != torch.float32 or weight.device != sem.device:
        raise RuntimeError("weight must be the float32 Conv2d weight [classes,S,1,1] (or [classes,S]) on sem's device")
    w2 = weight.detach().reshape(weight.shape[0], -1).contiguous()
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    return sem.detach().contiguous(), w2, b


class _LeafCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sem, labels, weight, bias, loss_weight, num_valid):
        lib = _lib.load()
        s, w2, b = _leaf_args(sem, weight, bias)
        if labels.shape != sem.shape[1:]:
            raise RuntimeError("labels must be [H,W]")
        lab = labels.to(torch.int32).contiguous()
        loss = torch.zeros((), dtype=torch.float32, device=s.device)
        grad = torch.empty_like(s)
        with torch.cuda.device(s.device):
            gw, gb = _run_leaf(lib, s, lab, w2, b, loss_weight, num_valid, loss, grad, False,
                               ctx.needs_input_grad[2] or (bias is not None and ctx.needs_input_grad[3]))
        ctx.wshape = weight.shape
        ctx.save_for_backward(grad, gw, gb)
        return loss

    @staticmethod
    def backward(ctx, g):
        grad, gw, gb = ctx.saved_tensors
        return (grad * g, None, None if gw is None else (gw * g).view(ctx.wshape), None if gb is None else gb * g,
                None, None)


def leaf_cross_entropy(sem: torch.Tensor, labels: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None,
                       loss_weight: float = 1.0, num_valid: int | None = None) -> torch.Tensor:
    """loss_weight * CrossEntropyLoss()(conv2d(sem[None], weight, bias)[0].view(classes, -1).T, labels.view(-1)) -- the
    leaf ("cross-level") loss of scripts/hierslam.py:975-984 / :1009-1016 with MLP_func = Conv2d(S, classes, 1) -- in two
    kernels that never materialise the [classes,H,W] logits.  Gradients flow to sem, weight and bias.
    labels: [H,W] integer (negative = ignored); num_valid: number of non-ignored pixels if the caller knows it (saves
    a host sync)."""
    return _LeafCE.apply(sem, labels, weight, bias, float(loss_weight), num_valid)


class _TreeLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sem, labels, weight, bias, level_sizes, level_weight, leaf_weight, num_valid, level_valid):
        lib = _lib.load()
        s, w2, b = _leaf_args(sem, weight, bias)
        nl = len(level_sizes)
        if labels.dim() != 3 or labels.shape[0] != nl + 1 or labels.shape[1:] != sem.shape[1:]:
            raise RuntimeError("labels must be [len(level_sizes) + 1, H, W]: one map per tree level, then the leaf map")
        if sum(level_sizes) > s.shape[0] or nl > 8:
            raise RuntimeError("level_sizes must sum to at most S and have at most 8 levels")
        lab = labels.to(torch.int32).contiguous()
        loss = torch.zeros((), dtype=torch.float32, device=s.device)
        grad = torch.zeros_like(s) if sum(level_sizes) < s.shape[0] else torch.empty_like(s)
        with torch.cuda.device(s.device):
            _run_hier(lib, s, lab[:nl], level_sizes, [level_weight] * nl, loss, grad, level_valid)
            gw, gb = _run_leaf(lib, s, lab[nl], w2, b, leaf_weight, num_valid, loss, grad, True,
                               ctx.needs_input_grad[2] or (bias is not None and ctx.needs_input_grad[3]))
        ctx.wshape = weight.shape
        ctx.save_for_backward(grad, gw, gb)
        return loss

    @staticmethod
    def backward(ctx, g):
        grad, gw, gb = ctx.saved_tensors
        return (grad * g, None, None if gw is None else (gw * g).view(ctx.wshape), None if gb is None else gb * g,
                None, None, None, None, None)


def tree_semantic_loss(sem: torch.Tensor, labels: torch.Tensor, level_sizes, weight: torch.Tensor,
                       bias: torch.Tensor | None = None, level_weight: float = 1.0, leaf_weight: float = 5.0,
                       num_valid: int | None = None, level_valid=None) -> torch.Tensor:
    """Hier-SLAM's whole semantic mapping loss (scripts/hierslam.py:955-984): level_weight * sum over the tree levels of
    the per-level cross-entropy + leaf_weight * the leaf cross-entropy behind the 1x1 convolution (weight_sem = [1, 5]).
    labels: [levels + 1, H, W] like curr_data['semantic_label_gt'] (last map = leaf labels).  One gradient image is
    produced: the leaf kernels accumulate onto what the level kernel wrote.  num_valid / level_valid: the numbers of
    non-ignored leaf / per-level pixels when known (no host sync then); int32 labels are used without conversion."""
    return _TreeLoss.apply(sem, labels, weight, bias, [int(v) for v in level_sizes], float(level_weight),
                           float(leaf_weight), num_valid, level_valid)


_WINDOW11 = None


def _window11():
    """the normalised 11-tap Gaussian (sigma 1.5) of utils/slam_external.py:55-57, rounded to float32 the same way"""
    global _WINDOW11
    if _WINDOW11 is None:
        from math import exp
        gauss = torch.tensor([exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)], dtype=torch.float32)
        _WINDOW11 = (ctypes.c_float * 11)(*(gauss / gauss.sum()).tolist())
    return _WINDOW11


class _L1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, l1_weight, ssim_weight):
        lib = _lib.load()
        if not pred.is_cuda:
            raise RuntimeError("l1_ssim_loss is CUDA-only (no CPU fallback)")
        if pred.dtype != torch.float32 or target.dtype != torch.float32 or pred.shape != target.shape or pred.dim() != 3:
            raise RuntimeError("pred and target must be float32 tensors of the same shape [C,H,W]")
        p, t = pred.detach().contiguous(), target.detach().contiguous()
        C, H, W = p.shape
        n = p.numel()
        loss = torch.full((), float(ssim_weight), dtype=torch.float32, device=p.device)
        scratch = torch.empty(3 * n, dtype=torch.float32, device=p.device)
        grad = torch.empty_like(p)
        with torch.cuda.device(p.device):
            _lib.check(lib.hs_l1_ssim(ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(t.data_ptr()), C, H, W, _window11(),
                                      float(l1_weight) / n, -float(ssim_weight) / n, ctypes.c_void_p(loss.data_ptr()),
                                      ctypes.c_void_p(scratch.data_ptr()), ctypes.c_void_p(grad.data_ptr()),
                                      _stream(p.device)), "hs_l1_ssim")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


def l1_ssim_loss(pred: torch.Tensor, target: torch.Tensor, l1_weight: float = 0.8, ssim_weight: float = 0.2) -> torch.Tensor:
    """l1_weight * |pred - target|.mean() + ssim_weight * (1 - calc_ssim(pred, target)) -- the colour loss of mapping
    (scripts/hierslam.py:936, utils/slam_external.py:55-97) -- value and gradient w.r.t. pred in two kernels instead of
    ten depthwise 11x11 convolutions.  pred / target: [C,H,W] float32 CUDA."""
    return _L1SSIM.apply(pred, target, float(l1_weight), float(ssim_weight))
