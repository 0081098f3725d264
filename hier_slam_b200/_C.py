"""Python host of the C ABI with the SAME five entry points, argument order and return tuples as the
reference's pybind module ``diff_gaussian_rasterization._C`` (reference: ext.cpp:15-23,
rasterize_points.cu:35-432), so the reference's own ``__init__.py`` could be run against it unchanged.

torch is used here only for what the task calls plumbing: device memory (tensors from the caching
allocator), the current stream and the device guard.  All arithmetic happens in libhsraster.so.
"""
from __future__ import annotations

import ctypes
import os
import threading
import weakref
from typing import Optional, Tuple

import torch

from . import _lib

# -- behaviour switches ------------------------------------------------------------------------------
# 'ref'   : reproduce the reference's observable semantic backward (quirk Q1):
#           semantic channels do not contribute to dL/dalpha.
# 'exact' : the mathematically intended gradient.
SEM_ALPHA_GRAD = os.environ.get("HS_SEM_ALPHA_GRAD", "ref")
NO_CULL = os.environ.get("HS_NO_CULL", "0") == "1"
BWD_SIMT = os.environ.get("HS_BWD_SIMT", "0") == "1"   # SIMT blend backward instead of the tensor-core one
SORT_GLOBAL = os.environ.get("HS_SORT_GLOBAL", "0") == "1"   # reference-style global radix sort instead of tile buckets
# Speculative binning of the drop-in path: from the second forward at an image size on, the binning buffer and the sort
# class are sized from the previous frame's counts (+25 %), the whole forward is enqueued without waiting, and the counts
# are read back while the blend kernel runs; the rare frame that did not fit is repeated synchronously.  `num_rendered`
# handed to autograd is then the buffer's capacity (it is opaque to callers: reference __init__.py:304-308).  Off: the
# forward waits for the counts before it sizes the buffer, like the reference (rasterizer_impl.cu:548), and returns the
# exact count.
SPECULATIVE = os.environ.get("HS_SPECULATIVE_BINNING", "1") == "1"


class exact_num_rendered:
    """with exact_num_rendered(): forwards of this thread wait for the instance count (no speculation), so that
    `num_rendered` and the state buffers' layout are the reference's (used by the parity tests that compare them)."""

    def __enter__(self):
        self.prev = getattr(_tls, "exact", False)
        _tls.exact = True

    def __exit__(self, *exc):
        _tls.exact = self.prev
        return False

_contig_cache: dict = {}   # id(tensor) -> (weakref(tensor), version, contiguous copy)


# -- sync-free binning ----------------------------------------------------------------------------------
class BinningCapacity:
    """Sizes for a forward without the `num_rendered` read-back (HS_ASYNC_BINNING, include/hs_raster.h): the binning buffer
    is allocated for `instances` (Gaussian, tile) pairs and the per-tile sort for lists of at most `longest_tile`
    entries; `num_rendered` handed to autograd is then the capacity.  Nothing synchronises, so the forward can be
    captured in a CUDA graph.  A frame that does not fit renders EMPTY and raises its overflow flag: check
    `overflowed()` (one host sync) before using results, and repeat synchronously with a larger capacity."""

    def __init__(self, instances: int, longest_tile: int):
        self.instances = max(int(instances), 1)
        self.longest_tile = max(1, min(int(longest_tile), 16384))
        self.infos: list = []      # int32[4] device views {instances, longest list, short lists, overflow}, one per forward

    @classmethod
    def from_info(cls, info: torch.Tensor, slack: float = 1.3, extra_instances: int = 65536,
                  extra_tile: int = 256) -> "BinningCapacity":
        """capacity with head-room from the counts of a forward (`binning_info` of its image buffer; host sync)"""
        r, longest = (int(v) for v in info[:2].tolist())
        return cls(int(r * slack) + extra_instances, int(longest * slack) + extra_tile)

    def overflowed(self) -> bool:
        if not self.infos:
            return False
        return bool(torch.stack([i[3] for i in self.infos]).any())

    def clear(self) -> None:
        self.infos.clear()


# capacity / recorder contexts are per THREAD: two rasterizer users on different threads (and streams) do not see each
# other's mode
_tls = threading.local()
_last_rendered: dict = {}   # (H, W, device) -> num_rendered of the previous forward at that image size (binning-buffer guess)


def _current_capacity() -> Optional[BinningCapacity]:
    return getattr(_tls, "capacity", None)


def _current_recorder() -> Optional[list]:
    return getattr(_tls, "recorder", None)


class async_binning:
    """with async_binning(capacity): every forward of THIS THREAD inside runs without a host sync (see BinningCapacity)."""

    def __init__(self, capacity: Optional[BinningCapacity]):
        self.capacity = capacity

    def __enter__(self):
        self.prev, _tls.capacity = _current_capacity(), self.capacity
        return self.capacity

    def __exit__(self, *exc):
        _tls.capacity = self.prev
        if self.capacity is not None and len(self.capacity.infos) > 4096:     # bound what a long-running caller accumulates
            del self.capacity.infos[:-4096]
        return False


class record_binning:
    """with record_binning() as infos: the int32[4] binning counts of every forward of this thread inside are appended to
    `infos` (used to size a BinningCapacity from synchronous renders)."""

    def __enter__(self):
        self.prev, _tls.recorder = _current_recorder(), []
        return _tls.recorder

    def __exit__(self, *exc):
        _tls.recorder = self.prev
        return False


def binning_info(imgBuffer: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """int32[4] view {num_rendered, longest tile list, short lists, capacity overflow} into a forward's image buffer"""
    off = _lib.load().hs_image_state_info_offset(int(H), int(W))
    return imgBuffer[off:off + 16].view(torch.int32)

# -- gradient sinks ------------------------------------------------------------------------------------
# A caller that keeps its parameters in one flat buffer (hier_slam_b200.mapping.FlatParams) can register, per
# parameter tensor, the gradient buffer that belongs to it.  When `colors_precomp` / `semantics_precomp` of a
# backward call ARE such parameters (same memory), the blend backward accumulates dL/dcolors and dL/dsemantics
# straight into the registered buffers -- they are pure atomic accumulations anyway -- and returns None for them,
# so autograd launches no AccumulateGrad add kernel and the call needs no [P, 3+S] zero fill.  With K keyframes per
# mapping iteration the K gradients sum up in place.
_grad_sinks: dict = {}     # data_ptr -> (weakref(param), grad buffer)


def register_grad_sink(param: torch.Tensor, grad: torch.Tensor) -> None:
    if not (param.is_cuda and grad.is_cuda and grad.dtype == torch.float32 and grad.is_contiguous()
            and grad.shape == param.shape):
        raise RuntimeError("gradient sink must be a contiguous float32 CUDA tensor of the parameter's shape")
    for k in [k for k, v in _grad_sinks.items() if v[0]() is None]:     # parameters that are gone release their buffers
        del _grad_sinks[k]
    _grad_sinks[param.data_ptr()] = (weakref.ref(param), grad)


def unregister_grad_sink(param: torch.Tensor) -> None:
    hit = _grad_sinks.get(param.data_ptr())
    if hit is not None and hit[0]() is param:
        del _grad_sinks[param.data_ptr()]


def clear_grad_sinks() -> None:
    _grad_sinks.clear()


def _sink_for(t: Optional[torch.Tensor]):
    """The registered gradient buffer if -- and only if -- `t` is the registered parameter itself and autograd will
    accumulate into exactly that buffer: a leaf that requires grad whose .grad IS the registered view.  Aliases of the
    parameter's memory (leaf.detach(), a no-grad copy of the leaf, a view) take the normal path, so e.g. tracking renders
    between two mapping steps cannot leak gradients into the mapping buffer."""
    if t is None or not _grad_sinks or t.numel() == 0 or not t.is_contiguous():
        return None
    hit = _grad_sinks.get(t.data_ptr())
    if hit is None:
        return None
    p = hit[0]()
    if p is None or p.data_ptr() != t.data_ptr() or p.shape != t.shape or hit[1].device != t.device:
        _grad_sinks.pop(t.data_ptr(), None)
        return None
    if not (t.requires_grad and t.is_leaf and t.grad is not None and t.grad.data_ptr() == hit[1].data_ptr()):
        return None
    return hit[1]


def _small_contig(t: torch.Tensor) -> torch.Tensor:
    """Contiguous float32 version of a small camera tensor, cached per tensor object + version
    (the [1,4,4] matrices of utils/recon_helpers.py:8-13 are transposed views and are reused every frame)."""
    if t.is_contiguous() and t.dtype == torch.float32:
        return t
    key = id(t)
    hit = _contig_cache.get(key)
    if hit is not None and hit[0]() is t and hit[1] == t._version:
        return hit[2]
    c = t.contiguous().float()
    if len(_contig_cache) > 256:
        for k in [k for k, v in _contig_cache.items() if v[0]() is None]:
            del _contig_cache[k]
    _contig_cache[key] = (weakref.ref(t), t._version, c)
    return c


def _ptr(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor, name: str, device) -> torch.Tensor:
    if t.numel() == 0:
        return t
    if t.device != device:
        raise RuntimeError(f"{name} must live on {device}, got {t.device}")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _camera(background, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width,
            campos, prefiltered, debug):
    keep = (_small_contig(viewmatrix), _small_contig(projmatrix), _small_contig(background),
            _small_contig(campos) if campos is not None and campos.numel() else None)
    cam = _lib.HsCamera(int(image_height), int(image_width), float(tan_fovx), float(tan_fovy), float(scale_modifier),
                        keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(),
                        keep[3].data_ptr() if keep[3] is not None else None, int(bool(prefiltered)), int(bool(debug)))
    return cam, keep


_BUILT_S = (16, 26, 32, 48, 64, 74, 102)   # semantic channel counts instantiated in libhsraster (hs_supports_semantic_channels)


def _padded_channels(S: int) -> int:
    """Kernels are instantiated for the channel counts Hier-SLAM ships (config.h:18: 16 / 26 / 74 / 102) and a few in
    between.  Any other S <= 102 (another label tree) runs on the next larger instantiation with zero-padded columns: the
    extra channels blend zeros, receive zero upstream gradients and are sliced away again.  S > 102 is rendered in
    several passes (`_chunks`)."""
    for b in _BUILT_S:
        if S <= b:
            return b
    raise RuntimeError(f"semantic channel count S={S} is not instantiated in libhsraster (built: 0,16,26,32,48,64,74,102; "
                       f"other values up to 102 are zero-padded)")


_CHUNK = 74     # channels per pass when S exceeds the widest instantiation


def _chunks(S: int):
    """[(first channel, channels, instantiated width)] of the passes that render S > 102 semantic channels (the reference
    supports any NUM_SEMANTIC by recompiling, cuda_rasterizer/config.h:18; e.g. a flat 550-class map).  Geometry, tile
    lists, transmittance and n_contrib do not depend on the semantic channels, so every pass blends over the SAME sorted
    lists (HS_REUSE_BINNING) and the backward -- linear in the upstream gradients -- is the sum of the per-pass backwards."""
    out, c0 = [], 0
    while c0 < S:
        n = min(_CHUNK, S - c0)
        out.append((c0, n, _padded_channels(n)))
        c0 += n
    return out


def _forward(background, means3D, colors, semantics, opacity, scales, rotations, scale_modifier, cov3D_precomp,
             viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered,
             debug, semantic: bool, extra_semantic_chunks=(), _speculate: bool = True):
    lib = _lib.load()
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")   # rasterize_points.cu:266-268
    if not means3D.is_cuda:
        raise RuntimeError("hier_slam_b200 rasterizer is CUDA-only (no CPU fallback): means3D must be a CUDA tensor")
    device = means3D.device
    P = means3D.size(0)
    H, W = int(image_height), int(image_width)
    S = int(semantics.size(1)) if (semantic and semantics.numel() > 0) else 0
    if semantic and semantics.numel() > 0 and (semantics.dim() != 2 or semantics.size(0) != P):
        raise RuntimeError("semantics_precomp must have dimensions (num_points, S)")
    if not lib.hs_supports_semantic_channels(S):
        raise RuntimeError(f"semantic channel count S={S} is not instantiated in libhsraster (built: 0,16,26,32,48,64,74,102)")
    use_sh = colors.numel() == 0
    if use_sh and P > 0:
        # spherical-harmonics colour path (reference forward.cu:20-71): sh is [P, M, 3]
        if sh is None or sh.numel() == 0 or sh.dim() != 3 or sh.size(0) != P or sh.size(2) != 3:
            raise RuntimeError("provide colors_precomp [P,3] or shs [P,M,3]")
        if not (0 <= int(degree) <= 3) or sh.size(1) < (int(degree) + 1) ** 2:
            raise RuntimeError(f"sh_degree {degree} needs 0 <= degree <= 3 and at least {(int(degree) + 1) ** 2} "
                               f"coefficients per Gaussian, got {sh.size(1)}")
        if campos is None or campos.numel() != 3:
            raise RuntimeError("campos [3] is required by the spherical-harmonics colour path")
    fopt = dict(dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        radii = torch.empty(P, dtype=torch.int32, device=device)
        byte = dict(dtype=torch.uint8, device=device)
        if P == 0:
            # reference: outputs are zero-filled and the rasterizer is not run (rasterize_points.cu:277-292)
            z = lambda c: torch.zeros(c, H, W, **fopt)
            empty = torch.empty(0, **byte)
            outs = (z(3), z(S) if semantic else None, z(1), z(1), z(1), z(1))
            return 0, outs, radii, empty, empty, empty
        means3D_c = _f32c(means3D, "means3D", device)
        colors_c = None if use_sh else _f32c(colors, "colors_precomp", device)
        sh_c = _f32c(sh, "shs", device) if use_sh else None
        M = int(sh.size(1)) if use_sh else 0
        sem_c = _f32c(semantics, "semantics_precomp", device) if S else None
        opac_c = _f32c(opacity, "opacities", device)
        scales_c = _f32c(scales, "scales", device)
        rot_c = _f32c(rotations, "rotations", device)
        cov_c = _f32c(cov3D_precomp, "cov3D_precomp", device)
        cam, keep = _camera(background, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, campos,
                            prefiltered, debug)
        # geometry arrays + the packed per-Gaussian records the blend kernel's TMA gather reads (sized for the widest pass)
        S_rows = max([S] + [int(c.size(1)) for c in extra_semantic_chunks])
        geom_bytes = lib.hs_geom_state_bytes_rows(P, S_rows)
        geomBuffer = torch.empty(geom_bytes, **byte)
        img_bytes = lib.hs_image_state_bytes(H, W)
        imgBuffer = torch.empty(img_bytes, **byte)
        cap = None if SORT_GLOBAL else _current_capacity()
        _recorder = _current_recorder()
        # Everything that does not depend on num_rendered is allocated BEFORE the read-back, and the binning buffer is
        # sized from the previous frame at this image size (+25 %; the map grows slowly from frame to frame), so the GPU idles only for the copy and two launches
        # between hs_forward_geometry's sync and the first kernel of hs_forward_render.
        out_color = torch.empty(3, H, W, **fopt)
        out_sem = torch.empty(S, H, W, **fopt) if semantic else None
        out_depth = torch.empty(1, H, W, **fopt)
        out_median = torch.empty(1, H, W, **fopt)
        out_opacity = torch.empty(1, H, W, **fopt)
        out_mask = None if semantic else torch.empty(1, H, W, **fopt)
        key = (H, W, device.index)
        prev = _last_rendered.get(key)             # (instances, longest tile list) of the previous forward at this size
        spec = (cap is None and prev is not None and SPECULATIVE and _speculate and not SORT_GLOBAL and not debug
                and not getattr(_tls, "exact", False) and prev[1] * 5 // 4 + 64 <= 16384
                and not torch.cuda.is_current_stream_capturing())
        if spec:     # capacities from the previous frame: nothing below waits for the device
            cap_inst, cap_tile = prev[0] * 5 // 4 + 4096, min(prev[1] * 5 // 4 + 64, 16384)
        guess = cap.instances if cap is not None else (cap_inst if spec else (prev[0] * 5 // 4 if prev else 0))
        binningBuffer = torch.empty(lib.hs_binning_state_bytes(guess), **byte) if guess > 0 else None
        R = ctypes.c_int(cap.instances if cap is not None else (cap_inst if spec else 0))
        hint = ctypes.c_int(cap.longest_tile if cap is not None else (cap_tile if spec else 0))
        gflags = _lib.HS_SORT_GLOBAL if SORT_GLOBAL else (_lib.HS_ASYNC_BINNING if cap is not None else 0)
        if spec:
            gflags = _lib.HS_ASYNC_BINNING | _lib.HS_DEFER_READBACK
        _lib.check(lib.hs_forward_geometry(ctypes.byref(cam), P, _ptr(means3D_c), _ptr(opac_c), _ptr(scales_c),
                                           _ptr(rot_c), _ptr(cov_c), _ptr(sh_c), int(degree), M, _ptr(radii),
                                           _ptr(geomBuffer), geom_bytes,
                                           _ptr(imgBuffer), img_bytes, gflags,
                                           ctypes.byref(R), ctypes.byref(hint), stream), "hs_forward_geometry")
        num_rendered = int(R.value)          # capacity mode: the capacity (the counts stay on the device)
        if cap is not None:
            cap.infos.append(binning_info(imgBuffer, H, W))
        elif not spec and not SORT_GLOBAL:
            # the synchronous call returned the instance count and (in the low bits of the hint) the longest tile list
            _last_rendered[key] = (num_rendered, int(hint.value) & 0x0fffffff)
        if _recorder is not None and not SORT_GLOBAL:
            _recorder.append(binning_info(imgBuffer, H, W))
        if binningBuffer is None or num_rendered > guess:
            binningBuffer = torch.empty(lib.hs_binning_state_bytes(num_rendered), **byte)
        bin_bytes = binningBuffer.numel()
        flags = _lib.HS_NO_CULL if NO_CULL else 0
        _lib.check(lib.hs_forward_render(ctypes.byref(cam), P, S, num_rendered, int(hint.value), _ptr(colors_c),
                                         _ptr(sem_c),
                                         _ptr(radii), _ptr(geomBuffer), geom_bytes, _ptr(binningBuffer), bin_bytes,
                                         _ptr(imgBuffer), img_bytes, _ptr(out_color), _ptr(out_sem), _ptr(out_depth),
                                         _ptr(out_median), _ptr(out_opacity), _ptr(out_mask), flags, stream),
                   "hs_forward_render")
        if spec:
            # the counts have been on their way since the tile scan; by now the blend kernel is queued behind them
            counts = (ctypes.c_int * 4)()
            _lib.check(lib.hs_forward_readback(counts), "hs_forward_readback")
            _last_rendered[key] = (int(counts[0]), int(counts[1]))
            if counts[3] != 0:       # did not fit the guessed capacity (rendered empty): once more, synchronously
                del keep
                return _forward(background, means3D, colors, semantics, opacity, scales, rotations, scale_modifier,
                                cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh,
                                degree, campos, prefiltered, debug, semantic, extra_semantic_chunks, _speculate=False)
        # more semantic channels than the widest instantiation: further blend passes over the same sorted lists
        extra_out = []
        for chunk in extra_semantic_chunks:
            chunk_c = _f32c(chunk, "semantics_precomp", device)
            Sc = int(chunk_c.size(1))
            o_sem = torch.empty(Sc, H, W, **fopt)
            _lib.check(lib.hs_forward_render(ctypes.byref(cam), P, Sc, num_rendered, int(hint.value), _ptr(colors_c),
                                             _ptr(chunk_c), _ptr(radii), _ptr(geomBuffer), geom_bytes, _ptr(binningBuffer), bin_bytes,
                                             _ptr(imgBuffer), img_bytes, _ptr(out_color), _ptr(o_sem), _ptr(out_depth),
                                             _ptr(out_median), _ptr(out_opacity), None, flags | _lib.HS_REUSE_BINNING,
                                             stream), "hs_forward_render")
            extra_out.append(o_sem)
        del keep
    if extra_semantic_chunks:
        return num_rendered, (out_color, out_sem, out_depth, out_median, out_opacity, out_mask), radii, geomBuffer, \
            binningBuffer, imgBuffer, extra_out
    return num_rendered, (out_color, out_sem, out_depth, out_median, out_opacity, out_mask), radii, geomBuffer, \
        binningBuffer, imgBuffer


def rasterize_gaussians_semantic(background, means3D, colors, semantics, opacity, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height,
                                 image_width, sh, degree, campos, prefiltered, debug):
    """reference: RasterizeGaussiansCUDA_semantic, rasterize_points.cu:240-336."""
    S = int(semantics.size(1)) if (semantics is not None and semantics.dim() == 2 and semantics.numel() > 0) else 0
    if S > _BUILT_S[-1] and means3D.size(0) > 0:
        pad = lambda t, w: t if t.size(1) == w else torch.nn.functional.pad(t, (0, w - t.size(1)))
        parts = [pad(semantics[:, c0:c0 + n], w) for c0, n, w in _chunks(S)]
        n, o, radii, gb, bb, ib, extra = _forward(background, means3D, colors, parts[0], opacity, scales, rotations,
                                                  scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                                  image_height, image_width, sh, degree, campos, prefiltered, debug, True,
                                                  extra_semantic_chunks=parts[1:])
        sem = torch.cat([img[:nc] for img, (c0, nc, w) in zip([o[1]] + extra, _chunks(S))], 0)
        return n, o[0], sem, o[2], o[3], o[4], radii, gb, bb, ib
    Sp = _padded_channels(S) if S > 0 else 0
    if Sp != S:
        semantics = torch.nn.functional.pad(semantics, (0, Sp - S))
    n, o, radii, gb, bb, ib = _forward(background, means3D, colors, semantics, opacity, scales, rotations,
                                       scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                       image_height, image_width, sh, degree, campos, prefiltered, debug, True)
    return n, o[0], (o[1][:S] if Sp != S else o[1]), o[2], o[3], o[4], radii, gb, bb, ib


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                        prefiltered, debug):
    """reference: RasterizeGaussiansCUDA, rasterize_points.cu:35-117 (returns out_mask before radii)."""
    n, o, radii, gb, bb, ib = _forward(background, means3D, colors, torch.empty(0), opacity, scales, rotations,
                                       scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                       image_height, image_width, sh, degree, campos, prefiltered, debug, False)
    return n, o[0], o[2], o[3], o[4], o[5], radii, gb, bb, ib


def _backward(background, means3D, radii, colors, semantics, scales, rotations, scale_modifier, cov3D_precomp,
              viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_semantic, dL_dout_depth,
              dL_dout_median_depth, dL_dout_final_opacity, sh, degree, campos, geomBuffer, R, binningBuffer,
              imageBuffer, debug, H: int, W: int, semantic: bool, pose_points=None):
    lib = _lib.load()
    device = means3D.device
    P = means3D.size(0)
    S = int(semantics.size(1)) if (semantic and semantics is not None and semantics.numel() > 0) else 0
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    use_sh = (colors is None or colors.numel() == 0) and M > 0
    fopt = dict(dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        # atomically accumulated outputs: ONE zero fill (minus what goes straight into registered gradient sinks)
        sink_c = _sink_for(colors) if not use_sh else None
        sink_s = _sink_for(semantics) if S else None
        nacc = 3 + 4 + 1 + (0 if sink_c is not None else 3) + (0 if sink_s is not None else S) + 1
        acc = torch.zeros(P * nacc, **fopt)
        o = 0

        def take(cols):
            nonlocal o
            v = acc[o:o + P * cols]
            o += P * cols
            return v
        dL_dconic = take(4).view(P, 2, 2)       # first: read back as float4 (needs 16-B alignment)
        dL_dsemantics = sink_s if sink_s is not None else take(S).view(P, S)
        dL_dmeans2D = take(3).view(P, 3)
        dL_dcolors = sink_c if sink_c is not None else take(3).view(P, 3)
        dL_dopacity = take(1).view(P, 1)
        dL_ddepths = take(1).view(P, 1)
        dL_dmeans3D = torch.empty(P, 3, **fopt)
        dL_dcov3D = torch.empty(P, 6, **fopt)
        have_scales = scales is not None and scales.numel() > 0
        dL_dscales = torch.empty(P, 3, **fopt) if have_scales else torch.zeros(P, 3, **fopt)
        dL_drotations = torch.empty(P, 4, **fopt) if have_scales else torch.zeros(P, 4, **fopt)
        dL_dsh = torch.empty(P, M, 3, **fopt) if use_sh else torch.zeros(P, M, 3, **fopt)
        dL_dpose = torch.zeros(3, 4, **fopt) if pose_points is not None else None
        if P != 0:
            stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            cam, keep = _camera(background, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, campos,
                                False, debug)
            g = lambda t, name: None if t is None else _f32c(t, name, device)
            gc, gs, gd, gm, go = (g(dL_dout_color, "dL_dout_color"), g(dL_dout_semantic, "dL_dout_semantic"),
                                  g(dL_dout_depth, "dL_dout_depth"), g(dL_dout_median_depth, "dL_dout_median_depth"),
                                  g(dL_dout_final_opacity, "dL_dout_final_opacity"))
            means3D_c = _f32c(means3D, "means3D", device)
            colors_c = None if use_sh else _f32c(colors, "colors_precomp", device)
            sh_c = _f32c(sh, "shs", device) if use_sh else None
            sem_c = _f32c(semantics, "semantics_precomp", device) if S else None
            scales_c = _f32c(scales, "scales", device) if have_scales else None
            rot_c = _f32c(rotations, "rotations", device) if have_scales else None
            cov_c = _f32c(cov3D_precomp, "cov3D_precomp", device) if cov3D_precomp is not None else None
            pose_c = _f32c(pose_points, "pose_points", device) if pose_points is not None else None
            flags = _lib.HS_SEM_ALPHA_EXACT if SEM_ALPHA_GRAD == "exact" else 0
            if BWD_SIMT:
                flags |= _lib.HS_BWD_SIMT
            _lib.check(lib.hs_backward(
                ctypes.byref(cam), P, S, int(R), _ptr(means3D_c), _ptr(radii), _ptr(colors_c), _ptr(sem_c),
                _ptr(scales_c), _ptr(rot_c), _ptr(cov_c), _ptr(sh_c), int(degree), int(M), _ptr(geomBuffer),
                _ptr(binningBuffer), _ptr(imageBuffer),
                _ptr(gc), _ptr(gs), _ptr(gd), _ptr(gm), _ptr(go), _ptr(dL_dmeans2D), _ptr(dL_dconic),
                _ptr(dL_dopacity), _ptr(dL_dcolors), _ptr(dL_dsemantics), _ptr(dL_ddepths), _ptr(dL_dmeans3D),
                _ptr(dL_dcov3D), _ptr(dL_dscales) if have_scales else None,
                _ptr(dL_drotations) if have_scales else None, _ptr(dL_dsh) if use_sh else None,
                _ptr(pose_c), _ptr(dL_dpose), flags, stream), "hs_backward")
            del keep
    if sink_c is not None:
        dL_dcolors = None       # already accumulated into the registered gradient buffer
    if sink_s is not None:
        dL_dsemantics = None
    if pose_points is not None:
        return (dL_dmeans2D, dL_dcolors, dL_dsemantics, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
                dL_drotations, dL_dpose)
    return dL_dmeans2D, dL_dcolors, dL_dsemantics, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


def _hw(image_buffer_hint, grads, H, W):
    for gten in grads:
        if gten is not None:
            return int(gten.size(1)), int(gten.size(2))
    if H is None or W is None:
        raise RuntimeError("image size unknown: all upstream gradients are None; pass image_height/image_width")
    return int(H), int(W)


def rasterize_gaussians_backward_semantic(background, means3D, radii, colors, semantics, scales, rotations,
                                          scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                          dL_dout_color, dL_dout_semantic, dL_dout_depth, dL_dout_median_depth,
                                          dL_dout_final_opacity, sh, degree, campos, geomBuffer, R, binningBuffer,
                                          imageBuffer, debug, image_height=None, image_width=None, pose_points=None):
    """reference: RasterizeGaussiansBackwardCUDA_semantic, rasterize_points.cu:339-432.  H and W are taken from
    dL_dout_color like the reference does (:369-370); the two optional trailing arguments supply them when the
    caller passes None for unmaterialised gradients.  pose_points (extension, hier_slam_b200.tracking): world-frame
    means; the tuple then ends with dL_dpose [3,4]."""
    H, W = _hw(imageBuffer, (dL_dout_color, dL_dout_semantic, dL_dout_depth, dL_dout_median_depth,
                             dL_dout_final_opacity), image_height, image_width)
    S = int(semantics.size(1)) if (semantics is not None and semantics.dim() == 2 and semantics.numel() > 0) else 0
    if S > _BUILT_S[-1] and means3D.size(0) > 0:
        # several passes over the same state (see _chunks): the first carries every non-semantic upstream gradient, the
        # others only their slice of dL/dout_semantic; geometry gradients add up, dL/dsemantics is concatenated
        total = None
        sem_grads = []
        for i, (c0, nc, w) in enumerate(_chunks(S)):
            sem_c = semantics[:, c0:c0 + nc]
            g_c = None if dL_dout_semantic is None else dL_dout_semantic[c0:c0 + nc]
            if w != nc:
                sem_c = torch.nn.functional.pad(sem_c, (0, w - nc))
                if g_c is not None:
                    g_c = torch.cat((g_c, g_c.new_zeros(w - nc, H, W)), 0)
            first = i == 0
            r = _backward(background, means3D, radii, colors, sem_c.contiguous(), scales, rotations, scale_modifier,
                          cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color if first else None,
                          None if g_c is None else g_c.contiguous(), dL_dout_depth if first else None,
                          dL_dout_median_depth if first else None, dL_dout_final_opacity if first else None, sh, degree,
                          campos, geomBuffer, R, binningBuffer, imageBuffer, debug, H, W, True, pose_points=pose_points)
            sem_grads.append(r[2][:, :nc])
            if total is None:
                total = list(r)
            elif g_c is not None:          # without an upstream gradient the pass contributes exact zeros
                for k in (0, 1, 3, 4, 5, 7, 8) + ((9,) if pose_points is not None else ()):
                    if total[k] is not None and r[k] is not None:
                        total[k] = total[k] + r[k]
        total[2] = torch.cat(sem_grads, 1)
        return tuple(total)
    Sp = _padded_channels(S) if S > 0 else 0
    if Sp != S:   # zero-padded instantiation (see _padded_channels): the forward state was produced with Sp channels
        semantics = torch.nn.functional.pad(semantics, (0, Sp - S))
        if dL_dout_semantic is not None:
            dL_dout_semantic = torch.cat((dL_dout_semantic, dL_dout_semantic.new_zeros(Sp - S, H, W)), 0)
    r = _backward(background, means3D, radii, colors, semantics, scales, rotations, scale_modifier, cov3D_precomp,
                  viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_semantic, dL_dout_depth,
                  dL_dout_median_depth, dL_dout_final_opacity, sh, degree, campos, geomBuffer, R, binningBuffer,
                  imageBuffer, debug, H, W, True, pose_points=pose_points)
    if Sp != S and r[2] is not None:
        r = r[:2] + (r[2][:, :S],) + r[3:]
    return r


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp,
                                 viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_depth,
                                 dL_dout_median_depth, dL_dout_final_opacity, sh, degree, campos, geomBuffer, R,
                                 binningBuffer, imageBuffer, debug, image_height=None, image_width=None, pose_points=None):
    """reference: RasterizeGaussiansBackwardCUDA, rasterize_points.cu:119-215.
    Returns (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)."""
    H, W = _hw(imageBuffer, (dL_dout_color, dL_dout_depth, dL_dout_median_depth, dL_dout_final_opacity),
               image_height, image_width)
    r = _backward(background, means3D, radii, colors, None, scales, rotations, scale_modifier, cov3D_precomp,
                  viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, None, dL_dout_depth,
                  dL_dout_median_depth, dL_dout_final_opacity, sh, degree, campos, geomBuffer, R, binningBuffer,
                  imageBuffer, debug, H, W, False, pose_points=pose_points)
    return (r[0], r[1], r[3], r[4], r[5], r[6], r[7], r[8]) + ((r[9],) if pose_points is not None else ())


def mark_visible(means3D, viewmatrix, projmatrix):
    """reference: markVisible, rasterize_points.cu:217-236."""
    lib = _lib.load()
    if not means3D.is_cuda:
        raise RuntimeError("hier_slam_b200 rasterizer is CUDA-only: means3D must be a CUDA tensor")
    P = means3D.size(0)
    present = torch.zeros(P, dtype=torch.bool, device=means3D.device)
    if P != 0:
        with torch.cuda.device(means3D.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(means3D.device).cuda_stream)
            m = _f32c(means3D, "means3D", means3D.device)
            v, pm = _small_contig(viewmatrix), _small_contig(projmatrix)
            _lib.check(lib.hs_mark_visible(P, _ptr(m), _ptr(v), _ptr(pm), _ptr(present), stream), "hs_mark_visible")
    return present


# ---- test / debugging aid: typed views into the opaque state buffers --------------------------------------
def state_views(P: int, H: int, W: int, R: int, geomBuffer, binningBuffer, imgBuffer):
    lib = _lib.load()
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    N = H * W

    def view(buf, off, n, dtype):
        esz = torch.empty((), dtype=dtype).element_size()
        return buf[off:off + n * esz].view(dtype)
    out = {}
    if P > 0:
        off = (ctypes.c_size_t * 5)()
        lib.hs_geom_state_layout(P, off)
        out["depths"] = view(geomBuffer, off[0], P, torch.float32)
        out["means2D"] = view(geomBuffer, off[1], 2 * P, torch.float32).view(P, 2)
        out["conic_opacity"] = view(geomBuffer, off[2], 4 * P, torch.float32).view(P, 4)
        out["tiles_touched"] = view(geomBuffer, off[3], P, torch.int32)
        out["point_offsets"] = view(geomBuffer, off[4], P, torch.int32)
    off = (ctypes.c_size_t * 3)()
    lib.hs_image_state_layout(H, W, off)
    out["final_T"] = view(imgBuffer, off[0], N, torch.float32)
    out["n_contrib"] = view(imgBuffer, off[1], N, torch.int32)
    out["ranges"] = view(imgBuffer, off[2], 2 * tiles, torch.int32).view(tiles, 2)
    if R > 0:
        off = (ctypes.c_size_t * 5)()
        lib.hs_binning_state_layout(R, off)
        out["point_list"] = view(binningBuffer, off[0], R, torch.int32)
        out["point_list_unsorted"] = view(binningBuffer, off[1], R, torch.int32)
        out["keys"] = view(binningBuffer, off[2], R, torch.int64)
        out["keys_unsorted"] = view(binningBuffer, off[3], R, torch.int64)
        out["strip_hits"] = view(binningBuffer, off[4], R, torch.uint8)
    return out
