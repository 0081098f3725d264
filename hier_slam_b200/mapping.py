"""Keyframe-parallel mapping across the GPUs of one box — a NEW capability: the reference
optimises ONE randomly chosen keyframe per mapping iteration on one GPU (scripts/hierslam.py:1986-2059).

Partitioning: one process per GPU (torchrun), the Gaussian set replicated, keyframe k of a K-keyframe batch
rendered by rank k mod G.  Every rank accumulates the gradients of its keyframes directly into ONE flat fp32
buffer (the leaves' ``.grad`` tensors are views into it, so autograd's accumulation *is* the pack step — there
is no gather/concat copy), then a single ``all_reduce(SUM)`` over NCCL / NVLink combines the ranks.  The
correctness criterion is: all-reduced gradient == sum of the K single-keyframe gradients (fp32 reassociation
only); tests/test_mapping_gloo.py checks it with world_size 2 on CPU (gloo), tests/test_gpu_parity.py on GPU.

Tracking (a sequential 40-100 iteration chain per frame) stays on one GPU: "replicas only".
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


class FlatParams:
    """Named fp32 leaves carved out of one flat buffer, with gradients carved out of one flat gradient buffer."""

    def __init__(self, tensors: Dict[str, torch.Tensor], direct_grads: bool = True):
        dev = next(iter(tensors.values())).device
        self._allocate({k: tuple(v.shape) for k, v in tensors.items()}, dev, direct_grads)
        for k, v in tensors.items():
            o, n = self.offsets[k], v.numel()
            self.flat[o:o + n].copy_(v.detach().reshape(-1).float())

    @classmethod
    def empty(cls, shapes: Dict[str, tuple], device, direct_grads: bool = True) -> "FlatParams":
        """Zero-initialised buffers and leaves for the given shapes (used by pruning: hier_slam_b200.optim)."""
        self = cls.__new__(cls)
        self._allocate({k: tuple(v) for k, v in shapes.items()}, torch.device(device), direct_grads)
        return self

    def _allocate(self, shapes: Dict[str, tuple], dev, direct_grads: bool) -> None:
        self.names: List[str] = list(shapes)
        self.shapes = dict(shapes)
        self.direct_grads = direct_grads
        sizes = [int(torch.Size(shapes[k]).numel()) for k in self.names]
        # 64-float (256-B) aligned segments so every view can be consumed with vector loads
        self.offsets = {}
        off = 0
        for k, n in zip(self.names, sizes):
            self.offsets[k] = off
            off += (n + 63) // 64 * 64
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.leaves: Dict[str, torch.Tensor] = {}
        for k, n in zip(self.names, sizes):
            o = self.offsets[k]
            leaf = self.flat[o:o + n].view(self.shapes[k]).requires_grad_(True)
            leaf.grad = self.flat_grad[o:o + n].view(self.shapes[k])
            self.leaves[k] = leaf
            if direct_grads and leaf.is_cuda and n > 0:
                # the rasterizer's backward accumulates dL/dcolors and dL/dsemantics straight into this buffer when a
                # leaf is handed to it unchanged (no autograd accumulate kernel, no [P, 3+S] zero fill per call)
                from . import _C
                _C.register_grad_sink(leaf, leaf.grad)

    @torch.no_grad()
    def appended(self, new_rows: Dict[str, torch.Tensor]) -> "FlatParams":
        """A new parameter set with `new_rows[name]` ([n, ...] each, same n) appended to every tensor -- what
        add_new_gaussians (scripts/hierslam.py:1307-1352: torch.cat per parameter) and cat_params_to_optimizer
        (utils/slam_external.py:124-139) do -- laid out in fresh flat buffers.  This set is left untouched; call
        release() on it when it is dropped."""
        n = {int(v.shape[0]) for v in new_rows.values()}
        if set(new_rows) != set(self.names) or len(n) != 1:
            raise RuntimeError("appended(): one [n, ...] tensor per parameter name, all with the same n")
        n = n.pop()
        shapes = {k: (self.shapes[k][0] + n,) + tuple(self.shapes[k][1:]) for k in self.names}
        new = FlatParams.empty(shapes, self.flat.device, self.direct_grads)
        for k in self.names:
            if tuple(new_rows[k].shape[1:]) != tuple(self.shapes[k][1:]):
                raise RuntimeError(f"appended(): {k} rows have shape {tuple(new_rows[k].shape[1:])}, expected {self.shapes[k][1:]}")
            old_n = int(torch.Size(self.shapes[k]).numel())
            add_n = int(new_rows[k].numel())
            o = new.offsets[k]
            new.flat[o:o + old_n].copy_(self.flat[self.offsets[k]:self.offsets[k] + old_n])
            new.flat[o + old_n:o + old_n + add_n].copy_(new_rows[k].detach().reshape(-1).float())
        return new

    def release(self) -> None:
        """Drop the gradient-sink registrations of the leaves (a replaced parameter set must not keep its buffers alive)."""
        if self.direct_grads:
            from . import _C
            for leaf in self.leaves.values():
                if leaf.is_cuda:
                    _C.unregister_grad_sink(leaf)

    def segment_ends(self) -> List[int]:
        """exclusive end (in floats, padding included) of every tensor's segment of the flat buffers, in name order"""
        return [self.offsets[k] for k in self.names[1:]] + [self.flat.numel()]

    def zero_grad(self) -> None:
        self.flat_grad.zero_()

    @torch.no_grad()
    def use_grad_buffer(self, buf: torch.Tensor) -> None:
        """Moves the flat gradient into `buf` (same size / dtype / device; e.g. a symmetric-memory allocation that the
        NVLink all-reduce kernel can address on every rank) and re-points every leaf's .grad and gradient sink at it."""
        if buf.shape != self.flat_grad.shape or buf.dtype != torch.float32 or buf.device != self.flat_grad.device:
            raise RuntimeError("use_grad_buffer: need a float32 buffer of the flat gradient's size on its device")
        buf.copy_(self.flat_grad)
        self.flat_grad = buf
        for k in self.names:
            leaf = self.leaves[k]
            o, n = self.offsets[k], leaf.numel()
            leaf.grad = buf[o:o + n].view(self.shapes[k])
            if self.direct_grads and leaf.is_cuda and n > 0:
                from . import _C
                _C.register_grad_sink(leaf, leaf.grad)

    def grad_bytes(self) -> int:
        return self.flat_grad.numel() * 4


def keyframes_of_rank(num_keyframes: int, rank: int, world_size: int) -> List[int]:
    """keyframe k -> rank k mod G."""
    return list(range(rank, num_keyframes, world_size))


ALLREDUCE_IMPL = "nccl all_reduce (torch.distributed)"    # what allreduce_gradients used last (bench.py reports it)


class SymmetricAllReduce:
    """The flat gradient buffer in NVLink symmetric memory + the peer tables of `hs_allreduce_sum`
    (hier_slam_b200/csrc/allreduce.cu: one kernel per rank; with NVSwitch multicast the reduction and the broadcast happen
    inside the switch through multimem.ld_reduce / multimem.st).  torch.distributed._symmetric_memory is used for the
    allocation, the rendezvous and the pointer tables only -- plumbing; the collective itself is this library's kernel."""

    def __init__(self, params: FlatParams, group=None, blocks: int = 64, multicast: bool = True):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = params.flat_grad.device
        n = params.flat_grad.numel()
        name = group.group_name
        try:
            symm.enable_symm_mem_for_group(name)            # needed by older releases; a no-op / deprecated later
        except Exception:
            pass
        self.buf = symm.empty(n, dtype=torch.float32, device=dev)
        self.hdl = symm.rendezvous(self.buf, name)
        self.pad = symm.empty(max(blocks * self.world, 1024), dtype=torch.int32, device=dev)
        self.pad.zero_()
        self.pad_hdl = symm.rendezvous(self.pad, name)
        torch.cuda.synchronize(dev)
        dist.barrier(group)                                  # every pad is zero before anybody signals into it
        self.blocks, self.epoch = blocks, 1
        self.multicast_ptr = int(self.hdl.multicast_ptr) if (multicast and self.hdl.has_multicast_support) else 0
        vp = ctypes.c_void_p
        self.peer_bufs = (vp * self.world)(*[vp(int(p)) for p in self.hdl.buffer_ptrs])
        self.peer_pads = (vp * self.world)(*[vp(int(p)) for p in self.pad_hdl.buffer_ptrs])
        params.use_grad_buffer(self.buf)
        self.count = n

    def __call__(self) -> None:
        import ctypes
        from . import _lib
        dev = self.buf.device
        _lib.check(_lib.load().hs_allreduce_sum(ctypes.c_void_p(self.multicast_ptr) if self.multicast_ptr else None,
                                                self.peer_bufs, self.peer_pads, self.rank, self.world, self.count, self.epoch,
                                                self.blocks, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "hs_allreduce_sum")
        self.epoch += 2

    @property
    def name(self) -> str:
        return ("hs_allreduce_sum: one kernel per rank over NVLink symmetric memory, " +
                ("NVSwitch multicast (multimem.ld_reduce + multimem.st)" if self.multicast_ptr else "peer loads / stores (two-shot)"))


def enable_symmetric_allreduce(params: FlatParams, group=None, blocks: int = 64, multicast: Optional[bool] = None):
    """Switches `allreduce_gradients(params)` to this library's NVLink kernel.  Returns the SymmetricAllReduce object, or
    None (with the reason in `.symm_error`) when symmetric memory cannot be set up -- NCCL's all_reduce is used then.
    multicast None = by world size: measured on 8 x B200 (tools/allreduce_bench.py, 48 MB): NVSwitch multicast 0.139 ms vs
    peer loads / stores 0.166 ms vs NCCL 0.223 ms; on 2 GPUs peer 0.095 ms vs multicast 0.144 ms vs NCCL 0.115 ms."""
    try:
        if multicast is None:
            multicast = dist.get_world_size(group) > 2
        params._symm = SymmetricAllReduce(params, group, blocks, multicast)
    except Exception as ex:          # no NVLink peer access / no fabric support: keep the NCCL collective
        params._symm = None
        params.symm_error = repr(ex)
    return params._symm


def allreduce_gradients(params: FlatParams, group=None) -> None:
    """ONE collective per mapping iteration on the flat gradient buffer (SUM, fp32)."""
    global ALLREDUCE_IMPL
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        symm = getattr(params, "_symm", None)
        if symm is not None:
            symm()
            ALLREDUCE_IMPL = symm.name
        else:
            dist.all_reduce(params.flat_grad, op=dist.ReduceOp.SUM, group=group)
            ALLREDUCE_IMPL = "nccl all_reduce (torch.distributed)"


def capacity_for(loss_fns, params: FlatParams, slack: float = 1.3):
    """A `_C.BinningCapacity` that fits every rasterizer call of the given keyframe losses (evaluated once, synchronously,
    without gradients) with `slack` head-room."""
    from . import _C
    if callable(loss_fns):
        loss_fns = [loss_fns]
    with torch.no_grad(), _C.record_binning() as infos:
        for f in loss_fns:
            f(params.leaves)
    if not infos:
        raise RuntimeError("capacity_for: the losses do not contain a rasterizer call")
    both = torch.stack(infos)[:, :2].max(0).values.tolist()
    return _C.BinningCapacity(int(both[0] * slack) + 65536, int(both[1] * slack) + 256)


_side_streams: Dict[int, List[torch.cuda.Stream]] = {}


def mapping_iteration(params: FlatParams, keyframe_losses: Sequence[Callable[[Dict[str, torch.Tensor]], torch.Tensor]],
                      rank: int = 0, world_size: int = 1, group=None, streams: int = 2, capacity=None) -> torch.Tensor:
    """One data-parallel mapping iteration.

    capacity (a `_C.BinningCapacity`, CUDA only): every forward of the iteration runs in capacity mode -- no
    `num_rendered` read-back, so the host never waits inside the iteration and a rank that renders a single keyframe
    (G = K) no longer exposes the launch latency behind each sync.  The overflow flags are checked ONCE, after the
    all-reduce has been enqueued (MAX over the ranks, so that all ranks agree); an iteration in which any keyframe
    outgrew the capacity is repeated synchronously.  `capacity_for(...)` sizes one from a synchronous render.
    Measured (tools/mapping_bench.py, one B200): with 8 local keyframes c4 goes 1391 -> 1461 keyframes/s; with ONE local
    keyframe it is slower (0.76 -> 0.89 ms per iteration), because the end-of-iteration flag read drains the pipeline
    that the synchronous path keeps full across iterations -- use it only when a rank renders several keyframes.

    keyframe_losses[k](leaves) renders keyframe k from the shared Gaussian leaves and returns its scalar loss.
    This rank evaluates keyframes k = rank, rank+G, ...; their gradients accumulate in params.flat_grad; the
    buffer is then all-reduced.  Returns this rank's summed loss (detached).

    With more than one local keyframe the keyframes alternate between `streams` CUDA streams: the forward of
    keyframe k+1 (including its num_rendered read-back, which waits on its own stream only) overlaps the backward of
    keyframe k, so the GPU is never idle behind the host sync and the tails of the blend kernels are filled.  The
    renders are independent; the gradient buffer is only ever accumulated into (atomics in the blend backward,
    autograd's stream-ordered AccumulateGrad for the rest)."""
    if capacity is not None:
        from . import _C
        capacity.clear()
        with _C.async_binning(capacity):
            total = mapping_iteration(params, keyframe_losses, rank, world_size, group, streams, None)
        flag = torch.stack([i[3] for i in capacity.infos]).max().to(torch.int32) if capacity.infos else \
            torch.zeros((), dtype=torch.int32, device=params.flat.device)
        if world_size > 1 and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)      # a single-rank iteration calls no collective
        if int(flag) == 0:                    # the iteration's one host sync, behind all of its GPU work
            return total
        return mapping_iteration(params, keyframe_losses, rank, world_size, group, streams, None)
    params.zero_grad()
    dev = params.flat.device
    total = torch.zeros((), device=dev)
    mine = keyframes_of_rank(len(keyframe_losses), rank, world_size)
    use_streams = dev.type == "cuda" and streams > 1 and len(mine) > 1
    if use_streams:
        pool = _side_streams.setdefault(dev.index, [])
        while len(pool) < streams:
            pool.append(torch.cuda.Stream(dev))
        main = torch.cuda.current_stream(dev)
        for st in pool[:streams]:
            st.wait_stream(main)              # zero_grad and every earlier use of the leaves
        parts = []
        # the leaves' AccumulateGrad nodes live on `main`; autograd orders them after the producing side stream (that
        # is the intended behaviour here, so its advisory warning about the stream mismatch is switched off)
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        try:
            for i, k in enumerate(mine):
                st = pool[i % streams]
                with torch.cuda.stream(st):
                    loss = keyframe_losses[k](params.leaves)
                    loss.backward()
                    parts.append(loss.detach())
        finally:
            if quiet is not None:
                quiet(True)
        for st in pool[:streams]:
            main.wait_stream(st)
        for part in parts:
            total = total + part
    else:
        for k in mine:
            loss = keyframe_losses[k](params.leaves)
            loss.backward()
            total = total + loss.detach()
    if world_size > 1:      # a single-rank iteration inside a larger job (e.g. a serial re-render for verification) must not
        allreduce_gradients(params, group)      # enter a collective that the other ranks do not call
    return total


def static_camera(raster_settings):
    """The same GaussianRasterizationSettings with contiguous float32 tensors of its own: what a captured graph may safely
    keep pointers to (the reference's setup_camera hands out transposed views, utils/recon_helpers.py:8-13)."""
    own = lambda t: t.detach().to(torch.float32).contiguous().clone()
    rs = raster_settings
    return rs._replace(bg=own(rs.bg), viewmatrix=own(rs.viewmatrix), projmatrix=own(rs.projmatrix), campos=own(rs.campos))


class GraphedMappingIteration:
    """One mapping iteration of Hier-SLAM (scripts/hierslam.py:1986-2059: pick a keyframe, get_loss_semantic_mlp,
    loss.backward(), optimizer.step(), zero_grad) as ONE CUDA-graph launch.

    `iteration()` must do the whole iteration on static inputs: zero the gradients, render from `params.leaves` (through
    GaussianRasterizer_semantic / PoseRasterizer_semantic), evaluate the loss with the kernels of hier_slam_b200.losses
    (pass num_valid / level_valid: no host sync), call backward, and step the optimisers (FlatAdam(device_step=True);
    torch optimisers with capturable=True).  It is run eagerly a few times (warm-up, and to size the binning capacity: the
    forward's only host read-back, `num_rendered`, is replaced by capacity-mode binning, `_C.BinningCapacity`), captured
    once and replayed.  What changes between iterations -- the keyframe's images, labels and pose -- lives in tensors
    that the caller overwrites in place before `replay()`.

    A replay whose frame outgrew the capacity renders empty; `overflowed()` (one host read, e.g. once per mapped frame)
    reports it, `recapture()` re-sizes.  Note: the optimiser also steps in the eager warm-up iterations."""

    def __init__(self, params: FlatParams, iteration: Callable[[], torch.Tensor], warmup: int = 2, slack: float = 1.3):
        from . import _C
        self.params, self.iteration, self.slack = params, iteration, slack
        self.graph = None
        self.capacity = None
        self.loss = None
        self._C = _C
        self._capture(warmup)

    def _capture(self, warmup: int, at_least=None) -> None:
        _C = self._C
        dev = self.params.flat.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with _C.record_binning() as infos:
                for _ in range(max(warmup, 1)):
                    self.iteration()
            if not infos:
                raise RuntimeError("GraphedMappingIteration: the iteration does not contain a rasterizer call")
            both = torch.stack(infos)[:, :2].max(0).values.tolist()
        torch.cuda.current_stream(dev).wait_stream(side)
        inst, longest = int(both[0] * self.slack) + 65536, int(both[1] * self.slack) + 256
        if at_least is not None:
            inst, longest = max(inst, at_least[0]), max(longest, at_least[1])
        self.capacity = _C.BinningCapacity(inst, longest)
        self.graph = torch.cuda.CUDAGraph()
        with _C.async_binning(self.capacity):
            with torch.cuda.graph(self.graph):
                self.loss = self.iteration()

    def replay(self) -> torch.Tensor:
        """one iteration; returns the (static) loss tensor of the captured graph -- read it only when needed"""
        self.graph.replay()
        return self.loss

    def overflowed(self) -> bool:
        return self.capacity.overflowed()

    def recapture(self, warmup: int = 1) -> None:
        """after an overflow (or when the number of Gaussians changed: build a new object then)"""
        need = torch.stack(self.capacity.infos)[:, :2].max(0).values.tolist() if self.capacity.infos else [0, 0]
        self._capture(warmup, (int(need[0] * self.slack) + 65536, int(need[1] * self.slack) + 256))
