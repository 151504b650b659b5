"""Autograd-aware collectives and pipeline point-to-point helpers.

These are the *unfused* communication primitives of the framework — the oracle
for, and the CPU/gloo fallback of, the fused sm_100a compute+collective kernels
in :mod:`quintnet_b200.ops`.  API parity: reference ``core/communication.py``
(``Send`` :46, ``Recv`` :129, ``pipeline_communicate`` :207,
``bidirectional_pipeline_communicate`` :299, ``All_Gather`` :374,
``All_Reduce`` :478, ``ReduceScatter`` :538).

Behavioural differences from the reference:

* the conjugate operator the reference lacks is provided
  (:class:`Copy_To_Parallel`: identity forward / all-reduce backward), so TP
  backward is mathematically complete (D2);
* collectives take a ``dim`` argument (the reference hard-codes the last dim),
  which is what sequence-parallel TP needs;
* pipeline p2p never calls ``torch.cuda.synchronize()`` (D11): receives are
  stream-ordered through the NCCL work handle, sends are left in flight and
  reaped lazily, so p2p overlaps with the stage's compute;
* a ``None`` group (axis of size 1) short-circuits every primitive.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

VERBOSE = os.environ.get("VERBOSE", "0") == "1"
STEP = 0


def group_size(group) -> int:
    if group is None:
        return 1
    return dist.get_world_size(group)


def group_rank(group) -> int:
    if group is None:
        return 0
    return dist.get_rank(group)


# --------------------------------------------------------------------------
# raw (non-autograd) helpers
# --------------------------------------------------------------------------
def all_reduce_(x: torch.Tensor, group, op=dist.ReduceOp.SUM) -> torch.Tensor:
    if group_size(group) > 1:
        dist.all_reduce(x, op=op, group=group)
    return x


def all_gather_cat(x: torch.Tensor, group, dim: int = -1) -> torch.Tensor:
    n = group_size(group)
    if n == 1:
        return x
    dim = dim % x.dim()
    x = x.contiguous()
    if dim == 0:
        out = x.new_empty((n * x.shape[0],) + tuple(x.shape[1:]))
        dist.all_gather_into_tensor(out, x, group=group)
        return out
    out = x.new_empty((n * x.shape[0],) + tuple(x.shape[1:]))
    dist.all_gather_into_tensor(out, x, group=group)
    # [n*d0, ...] -> n chunks concatenated along `dim`
    return torch.cat(out.chunk(n, dim=0), dim=dim)


def reduce_scatter_chunk(x: torch.Tensor, group, dim: int = -1) -> torch.Tensor:
    n = group_size(group)
    if n == 1:
        return x
    dim = dim % x.dim()
    if x.shape[dim] % n != 0:
        raise ValueError(f"dim {dim} of size {x.shape[dim]} is not divisible by group size {n}")
    if dim != 0:
        x = torch.cat([c for c in x.chunk(n, dim=dim)], dim=0)   # chunk r -> rows of rank r
    x = x.contiguous()
    out = x.new_empty((x.shape[0] // n,) + tuple(x.shape[1:]))
    _reduce_scatter_tensor(out, x, group)
    return out


_RS_UNSUPPORTED = set()


def _reduce_scatter_tensor(out: torch.Tensor, x: torch.Tensor, group) -> None:
    """reduce_scatter_tensor with an all-reduce+slice fallback for backends
    (older gloo) that lack the primitive."""
    backend = dist.get_backend(group)
    if backend not in _RS_UNSUPPORTED:
        try:
            dist.reduce_scatter_tensor(out, x, group=group)
            return
        except (RuntimeError, NotImplementedError):
            _RS_UNSUPPORTED.add(backend)
    full = x.clone()
    dist.all_reduce(full, group=group)
    n, r = dist.get_world_size(group), dist.get_rank(group)
    out.copy_(full.chunk(n, dim=0)[r])


def _sum_all_reduce(x: torch.Tensor, group) -> torch.Tensor:
    """Out-of-place SUM all-reduce.  Small CUDA messages (the latency-bound TP activations of the ViT path: <= 512 KB)
    go through the one-launch NVLS / P2P kernel over symmetric memory; everything else through the backend."""
    if x.is_cuda:
        from ..ops import symm
        if symm.oneshot_ok(x, group):
            return symm.all_reduce_small(x, group)
    out = x.contiguous().clone()  # never alias the producer's buffer
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


# --------------------------------------------------------------------------
# autograd collectives
# --------------------------------------------------------------------------
class All_Reduce(torch.autograd.Function):
    """SUM all-reduce forward, identity backward (Megatron's *g*)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, group=None):
        if group_size(group) == 1:
            return x
        return _sum_all_reduce(x, group)

    @staticmethod
    def backward(ctx, grad):
        return grad, None


class Copy_To_Parallel(torch.autograd.Function):
    """Identity forward, SUM all-reduce backward (Megatron's *f*).

    Missing from the reference: without it the input gradient of
    a column-parallel layer only holds this rank's partial sum.
    """

    @staticmethod
    def forward(ctx, x: torch.Tensor, group=None):
        ctx.group = group
        return x

    @staticmethod
    def backward(ctx, grad):
        if group_size(ctx.group) == 1:
            return grad, None
        return _sum_all_reduce(grad, ctx.group), None


class All_Gather(torch.autograd.Function):
    """All-gather + concat along ``dim``.

    ``mode='slice'``: backward keeps the local chunk (valid when the gathered
    tensor is consumed identically on every rank, gradient already replicated).
    ``mode='reduce_scatter'``: backward reduce-scatters the gradient (valid when
    each rank consumed the gathered tensor differently, e.g. sequence parallel).
    """

    @staticmethod
    def forward(ctx, x: torch.Tensor, group=None, mode: str = "slice", dim: int = -1):
        if mode not in ("slice", "reduce_scatter"):
            raise ValueError(f"unknown All_Gather mode {mode!r}")
        ctx.group, ctx.mode, ctx.dim = group, mode, dim % x.dim()
        return all_gather_cat(x, group, dim)

    @staticmethod
    def backward(ctx, grad):
        n = group_size(ctx.group)
        if n == 1:
            return grad, None, None, None
        if ctx.mode == "slice":
            g = grad.chunk(n, dim=ctx.dim)[group_rank(ctx.group)].contiguous()
        else:
            g = reduce_scatter_chunk(grad, ctx.group, ctx.dim)
        return g, None, None, None


class ReduceScatter(torch.autograd.Function):
    """Reduce-scatter along ``dim`` forward, all-gather backward."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, group=None, dim: int = -1):
        ctx.group, ctx.dim = group, dim % x.dim()
        return reduce_scatter_chunk(x, group, dim)

    @staticmethod
    def backward(ctx, grad):
        return all_gather_cat(grad, ctx.group, ctx.dim), None, None


class Scatter_To_Parallel(torch.autograd.Function):
    """Keep the local chunk along ``dim`` forward, all-gather backward."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, group=None, dim: int = -1):
        ctx.group, ctx.dim = group, dim % x.dim()
        n = group_size(group)
        if n == 1:
            return x
        return x.chunk(n, dim=ctx.dim)[group_rank(group)].contiguous()

    @staticmethod
    def backward(ctx, grad):
        return all_gather_cat(grad, ctx.group, ctx.dim), None, None


# --------------------------------------------------------------------------
# autograd point-to-point (exported for API parity; the schedules use the
# pipeline helpers below)
# --------------------------------------------------------------------------
def _send_with_header(t: torch.Tensor, dst: int, group) -> None:
    t = t.contiguous()
    header = torch.tensor([t.dim()] + list(t.shape) + [0] * (8 - t.dim()), dtype=torch.long,
                          device=t.device)
    dist.send(header, dst=dst, group=group)
    dist.send(t, dst=dst, group=group)


def _recv_with_header(src: int, group, device, dtype) -> torch.Tensor:
    header = torch.zeros(9, dtype=torch.long, device=device)
    dist.recv(header, src=src, group=group)
    h = header.tolist()
    shape = h[1:1 + h[0]]
    t = torch.empty(shape, device=device, dtype=dtype)
    dist.recv(t, src=src, group=group)
    return t


class Send(torch.autograd.Function):
    """Send ``x`` to global rank ``dst``; backward receives its gradient from there."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, dst: int, group=None):
        ctx.dst, ctx.group = dst, group
        ctx.meta = (x.shape, x.dtype, x.device)
        _send_with_header(x, dst, group)
        return x

    @staticmethod
    def backward(ctx, grad):
        shape, dtype, device = ctx.meta
        g = torch.empty(shape, dtype=dtype, device=device)
        dist.recv(g, src=ctx.dst, group=ctx.group)
        return g, None, None


class Recv(torch.autograd.Function):
    """Receive a tensor from global rank ``src``; backward sends the gradient back."""

    @staticmethod
    def forward(ctx, placeholder: torch.Tensor, src: int, group=None, dtype=None):
        ctx.src, ctx.group = src, group
        return _recv_with_header(src, group, placeholder.device, dtype or placeholder.dtype)

    @staticmethod
    def backward(ctx, grad):
        dist.send(grad.contiguous(), dst=ctx.src, group=ctx.group)
        return None, None, None, None


# --------------------------------------------------------------------------
# pipeline p2p
# --------------------------------------------------------------------------
class _InFlight:
    """Sends that have been issued but whose buffers must stay alive."""

    def __init__(self):
        self.items: List[Tuple[object, torch.Tensor]] = []

    def add(self, works, tensor):
        for w in works:
            self.items.append((w, tensor))
        # bound the backlog: reap everything but the newest few
        while len(self.items) > 8:
            w, _ = self.items.pop(0)
            w.wait()

    def drain(self):
        for w, _ in self.items:
            w.wait()
        self.items.clear()


_inflight = _InFlight()


def drain_pipeline_sends() -> None:
    """Wait (stream-ordered on CUDA) for all outstanding pipeline sends."""
    _inflight.drain()


def _peer(pp_group, pp_rank: int, delta: int) -> int:
    return dist.get_global_rank(pp_group, pp_rank + delta)


# Direction channels.  NCCL executes the p2p operations of ONE communicator in issue order on one stream, so a receive
# posted behind a send that still waits for its producer kernel cannot make progress.  With a dedicated communicator per
# direction (activations travel on the forward channel, gradients on the backward channel) a pre-posted receive of the
# NEXT micro-batch's activations completes while the current backward is still computing.  Registered per pp group by
# ``coordinators.compose`` (``DeviceMesh.get_channel_groups``); without a registration both directions use the pp group.
_channels: dict = {}


def register_pipeline_channels(pp_group, fwd_group, bwd_group) -> None:
    _channels[id(pp_group)] = (fwd_group, bwd_group)


def _channel(pp_group, forward: bool):
    ch = _channels.get(id(pp_group))
    if ch is None:
        return pp_group
    return ch[0] if forward else ch[1]


class PendingRecv:
    """A receive that has been POSTED (its buffer is being filled by the communication stream) but not yet waited for."""

    def __init__(self, buf: Optional[torch.Tensor], works, needs_grad: bool):
        self.buf, self.works, self.needs_grad = buf, works, needs_grad

    def wait(self) -> Optional[torch.Tensor]:
        if self.buf is None:
            return None
        for w in self.works:
            w.wait()                       # stream-ordered on CUDA: the compute stream waits, the host does not
        self.works = []
        if self.needs_grad:
            self.buf.requires_grad_(True)
        return self.buf


def post_recv(direction: str, pp_group, pp_rank: int, device, dtype, shapes, is_first_stage: bool = False,
              is_last_stage: bool = False) -> PendingRecv:
    """Post the receive of the next activation (``direction='forward'``, from the previous stage) or gradient
    (``'backward'``, from the next stage) and return immediately; ``.wait()`` when the tensor is consumed."""
    fwd = direction == "forward"
    if (fwd and is_first_stage) or (not fwd and is_last_stage):
        return PendingRecv(None, [], False)
    buf = torch.empty(tuple(shapes), device=device, dtype=dtype)
    g = _channel(pp_group, fwd)
    works = _launch([dist.P2POp(dist.irecv, buf, _peer(pp_group, pp_rank, -1 if fwd else +1), g)])
    return PendingRecv(buf, works, fwd)


def post_send(direction: str, pp_group, pp_rank: int, tensor: torch.Tensor, is_first_stage: bool = False,
              is_last_stage: bool = False) -> None:
    """Asynchronous send of an activation to the next stage (``'forward'``) or a gradient to the previous one."""
    fwd = direction == "forward"
    if (fwd and is_last_stage) or (not fwd and is_first_stage):
        return
    t = tensor.detach().contiguous()
    g = _channel(pp_group, fwd)
    _inflight.add(_launch([dist.P2POp(dist.isend, t, _peer(pp_group, pp_rank, +1 if fwd else -1), g)]), t)


def _launch(ops):
    return dist.batch_isend_irecv(ops) if ops else []


def pipeline_communicate(
    operation: str,
    pp_group,
    pp_rank: int,
    device,
    dtype,
    is_first_stage: bool = False,
    is_last_stage: bool = False,
    tensor: Optional[torch.Tensor] = None,
    shapes: Optional[Sequence[int]] = None,
):
    """One-directional stage-to-stage transfer.

    ``operation`` is one of ``recv_forward | send_forward | recv_backward |
    send_backward``.  Receives return a fresh tensor that is safe to use on the
    current stream; sends return ``None`` and complete asynchronously.
    """
    global STEP
    if operation == "recv_forward":
        if is_first_stage:
            return None
        out = post_recv("forward", pp_group, pp_rank, device, dtype, shapes).wait()
    elif operation == "recv_backward":
        if is_last_stage:
            return None
        out = post_recv("backward", pp_group, pp_rank, device, dtype, shapes).wait()
    elif operation == "send_forward":
        if is_last_stage:
            return None
        post_send("forward", pp_group, pp_rank, tensor)
        out = None
    elif operation == "send_backward":
        if is_first_stage:
            return None
        post_send("backward", pp_group, pp_rank, tensor)
        out = None
    else:
        raise ValueError(f"unknown pipeline operation {operation!r}")
    if VERBOSE:
        STEP += 1
        print(f"[p2p {STEP}] rank {dist.get_rank()} {operation}", flush=True)
    return out


def bidirectional_pipeline_communicate(
    operation: str,
    pp_group,
    pp_rank: int,
    send_tensor: Optional[torch.Tensor],
    recv_shapes: Sequence[int],
    device,
    dtype,
    is_first_stage: bool = False,
    is_last_stage: bool = False,
):
    """Batched send+recv with the same neighbour (1F1B steady state).

    ``send_fwd_recv_bwd`` talks to the next stage, ``send_bwd_recv_fwd`` to the
    previous one.  Returns the received tensor (``None`` at the pipeline edge).
    """
    global STEP
    if operation == "send_fwd_recv_bwd":
        if is_last_stage:
            return None
        peer, recv_needs_grad = _peer(pp_group, pp_rank, +1), False
    elif operation == "send_bwd_recv_fwd":
        if is_first_stage:
            return None
        peer, recv_needs_grad = _peer(pp_group, pp_rank, -1), True
    else:
        raise ValueError(f"unknown pipeline operation {operation!r}")
    send_fwd = operation == "send_fwd_recv_bwd"
    if id(pp_group) in _channels:
        # direction channels: the send stays in flight on its own communicator, only the receive is waited for
        post_send("forward" if send_fwd else "backward", pp_group, pp_rank, send_tensor)
        recv = post_recv("backward" if send_fwd else "forward", pp_group, pp_rank, device, dtype, recv_shapes).wait()
    else:
        recv = torch.empty(tuple(recv_shapes), device=device, dtype=dtype)
        t = send_tensor.detach().contiguous()
        works = _launch([
            dist.P2POp(dist.isend, t, peer, pp_group),
            dist.P2POp(dist.irecv, recv, peer, pp_group),
        ])
        # with batch_isend_irecv the works complete together; waiting is stream-ordered on CUDA
        for w in works:
            w.wait()
        if recv_needs_grad:
            recv.requires_grad_(True)
    if VERBOSE:
        STEP += 1
        print(f"[p2p {STEP}] rank {dist.get_rank()} {operation} peer={peer}", flush=True)
    return recv
