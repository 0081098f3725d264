"""Parameter maintenance on the flat buffers of `hier_slam_b200.mapping.FlatParams`.

`FlatAdam` is torch.optim.Adam as Hier-SLAM configures it (scripts/hierslam.py:411-417: one parameter group per named
tensor with its own learning rate, `eps=1e-15` in mapping) evaluated by ONE kernel over the flat parameter / gradient /
moment buffers (hs_adam_step) instead of seven multi-tensor kernels.  `FlatAdam.prune` is `remove_points`
(utils/slam_external.py:142-164): the kept rows of every parameter and of both Adam moments are gathered by three
kernels behind one keep-mask scan instead of one `tensor[to_keep]` (nonzero + host sync + gather) per tensor.
CUDA only; there is no CPU fallback."""
from __future__ import annotations

import ctypes
from typing import Dict

import torch

from . import _lib
from .mapping import FlatParams


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class FlatAdam:
    """device_step=True keeps the step count on the device (hs_adam_step_device): `step()` then contains no host value
    that changes from call to call, so it can be captured in a CUDA graph together with the render and the losses
    (hier_slam_b200.mapping.GraphedMappingIteration)."""

    def __init__(self, params: FlatParams, lrs: Dict[str, float], betas=(0.9, 0.999), eps: float = 1e-8,
                 device_step: bool = False):
        if not params.flat.is_cuda:
            raise RuntimeError("FlatAdam is CUDA-only (no CPU fallback)")
        missing = [k for k in params.names if k not in lrs]
        if missing:
            raise KeyError(f"no learning rate for {missing}")
        if len(params.names) > 16:
            raise RuntimeError("at most 16 parameter tensors per flat buffer")
        self.params = params
        self.lrs = {k: float(lrs[k]) for k in params.names}
        self.betas = (float(betas[0]), float(betas[1]))
        self.eps = float(eps)
        self.step_count = 0
        self.exp_avg = torch.zeros_like(params.flat)
        self.exp_avg_sq = torch.zeros_like(params.flat)
        self.device_step = bool(device_step)
        if self.device_step:
            self.step_dev = torch.zeros(1, dtype=torch.int32, device=params.flat.device)
            self.scalars_dev = torch.zeros(32, dtype=torch.float32, device=params.flat.device)

    def zero_grad(self) -> None:
        self.params.zero_grad()

    def state(self, name: str):
        """(exp_avg, exp_avg_sq) views of one parameter tensor, shaped like it"""
        p = self.params
        o, n = p.offsets[name], int(torch.Size(p.shapes[name]).numel())
        return self.exp_avg[o:o + n].view(p.shapes[name]), self.exp_avg_sq[o:o + n].view(p.shapes[name])

    @torch.no_grad()
    def step(self) -> None:
        lib = _lib.load()
        p = self.params
        self.step_count += 1
        n_seg = len(p.names)
        ends = (ctypes.c_ulonglong * n_seg)(*p.segment_ends())
        lrs = (ctypes.c_double * n_seg)(*[self.lrs[k] for k in p.names])
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(p.flat.device):
            if self.device_step:
                _lib.check(lib.hs_adam_step_device(vp(p.flat), vp(p.flat_grad), vp(self.exp_avg), vp(self.exp_avg_sq),
                                                   p.flat.numel(), n_seg, ends, lrs, self.betas[0], self.betas[1], self.eps,
                                                   vp(self.step_dev), vp(self.scalars_dev), _stream(p.flat.device)),
                           "hs_adam_step_device")
                return
            _lib.check(lib.hs_adam_step(vp(p.flat), vp(p.flat_grad), vp(self.exp_avg), vp(self.exp_avg_sq), p.flat.numel(),
                                        n_seg, ends, lrs, self.betas[0], self.betas[1], self.eps, self.step_count,
                                        _stream(p.flat.device)), "hs_adam_step")

    @torch.no_grad()
    def append(self, new_rows: Dict[str, torch.Tensor]) -> FlatParams:
        """cat_params_to_optimizer (utils/slam_external.py:124-139): new Gaussians join every parameter tensor, their Adam
        moments start at zero, the moments of the existing rows are kept.  Returns (and installs) the new FlatParams."""
        old = self.params
        new = old.appended(new_rows)
        new_m, new_v = torch.zeros_like(new.flat), torch.zeros_like(new.flat)
        for k in old.names:
            n = int(torch.Size(old.shapes[k]).numel())
            for src, dst in ((self.exp_avg, new_m), (self.exp_avg_sq, new_v)):
                dst[new.offsets[k]:new.offsets[k] + n].copy_(src[old.offsets[k]:old.offsets[k] + n])
        old.release()
        self.params, self.exp_avg, self.exp_avg_sq = new, new_m, new_v
        return new

    @torch.no_grad()
    def prune(self, keep: torch.Tensor) -> FlatParams:
        """Keep the rows (Gaussians) where `keep` is true in every parameter tensor and in both Adam moments, in their
        original order; returns (and installs) the new FlatParams.  One host sync (the number of kept rows)."""
        lib = _lib.load()
        old = self.params
        P = int(keep.numel())
        bad = [k for k in old.names if len(old.shapes[k]) == 0 or old.shapes[k][0] != P]
        if bad:
            raise RuntimeError(f"prune: {bad} do not have one row per entry of the keep mask ({P})")
        dev = old.flat.device
        keep_u8 = keep.to(device=dev, dtype=torch.bool).contiguous().view(torch.uint8)
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            scratch = torch.empty(max(int(lib.hs_compact_scratch_bytes(P)), 4), dtype=torch.uint8, device=dev)
            _lib.check(lib.hs_compact_plan(vp(keep_u8), P, vp(scratch), _stream(dev)), "hs_compact_plan")
            rows = int(scratch.view(torch.int32)[(P + 1023) // 1024]) if P > 0 else 0
            new = FlatParams.empty({k: (rows,) + tuple(old.shapes[k][1:]) for k in old.names}, dev, old.direct_grads)
            n_seg = len(old.names)
            width = [int(torch.Size(old.shapes[k][1:]).numel()) for k in old.names]
            src_off = (ctypes.c_ulonglong * n_seg)(*[old.offsets[k] for k in old.names])
            dst_off = (ctypes.c_ulonglong * n_seg)(*[new.offsets[k] for k in old.names])
            widths = (ctypes.c_int * n_seg)(*width)
            new_m, new_v = torch.zeros_like(new.flat), torch.zeros_like(new.flat)
            if rows > 0:
                for src, dst in ((old.flat, new.flat), (self.exp_avg, new_m), (self.exp_avg_sq, new_v)):
                    _lib.check(lib.hs_compact_gather(vp(src), vp(dst), vp(scratch), P, rows, n_seg, src_off, dst_off,
                                                     widths, _stream(dev)), "hs_compact_gather")
        old.release()
        self.params, self.exp_avg, self.exp_avg_sq = new, new_m, new_v
        return new
