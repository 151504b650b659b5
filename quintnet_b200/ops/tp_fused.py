"""Fused tensor-parallel GEMM + collective ops (K3 / K4) over NVLink symmetric memory.

``allgather_linear``      Y[M, N/tp]   = act( AG(X_local)[M, K] · W_localᵀ + b )        (ColumnParallel forward)
``linear_reduce_scatter`` Y_local[M/tp, N] = RS( X[M, K/tp] · W_localᵀ (+ b on its owner) )  (RowParallel forward)

Each is ONE launch of the tcgen05 GEMM kernel (``csrc/gemm.cu``, ``FUSE_AG`` / ``FUSE_RS``) that moves the
activations itself with TMA bulk copies on peer-mapped addresses: a one-thread copy pump per CTA streams the local
shard into the peers' gather buffers, or the epilogue bulk-stores partial slabs into the owner's staging buffer,
synchronised by release/acquire flags — no NCCL call.  The backward of
one op is the other op (plus local wgrad GEMMs), so a Megatron sequence-parallel block issues exactly four
fused GEMM+collective launches per direction.

Reference call sites these replace: ``ColumnParallelLinear.forward`` (``parallelism/tensor_parallel/layers.py:101-131``,
GEMM then optional ``All_Gather``) and ``RowParallelLinear.forward`` (``layers.py:198-221``, GEMM then
``All_Reduce`` — collective C3 of SURVEY §2.9, ``core/communication.py:515``).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import native, symm
from .cuda_ops import EPI_BIAS, EPI_BIAS_GELU, EPI_MUL_AUX, EPI_NONE, bias_grad, colsum, gemm, wgrad

_P, _I, _L, _U32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_uint
native.declare("qn_gemm_ag_bf16", [_P, _P, _P, _I, _I, _U32, _P, _P, _I, _I, _I, _L, _L, _I, _I, _P, _P, _L, _I, _I, _P, _P])
native.declare("qn_gemm_rs_bf16", [_P, _P, _P, _P, _P, _I, _I, _U32, _I, _I, _I, _L, _L, _L, _I, _P, _I, _I, _P])

_FLAG_BYTES = 1 << 20


class _Workspace:
    """Per-TP-group symmetric scratch: double-buffered gather / staging areas + monotonically increasing flags."""

    def __init__(self, group, device):
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.gather = [None, None]
        self.stage = [None, None]
        self.cap = 0
        self.ag_flags = symm.alloc_bytes(_FLAG_BYTES, device, group)
        self.rs_flags = symm.alloc_bytes(_FLAG_BYTES, device, group)
        self.ag_flags.tensor.zero_()
        self.rs_flags.tensor.zero_()
        self.seq_ag = 0
        self.seq_rs = 0
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def ensure(self, nbytes: int) -> None:
        if nbytes <= self.cap:
            return
        cap = max(nbytes, 32 << 20)
        for i in range(2):
            self.gather[i] = symm.alloc_bytes(cap, self.device, self.group)
            self.stage[i] = symm.alloc_bytes(cap, self.device, self.group)
        self.cap = cap
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    @staticmethod
    def ptrs(buf: symm.SymmBuffer):
        return (ctypes.c_void_p * buf.world)(*buf.peer_ptrs)


_workspaces: Dict[int, _Workspace] = {}
_unavailable: Optional[str] = None


def _ws(group, device) -> _Workspace:
    key = id(group)
    if key not in _workspaces:
        _workspaces[key] = _Workspace(group, device)
    return _workspaces[key]


def available(group) -> bool:
    return symm.available(group) and _unavailable is None


def can_fuse(rows_total: int, group, *dims) -> bool:
    """``rows_total`` = M of the (gathered) token matrix; dims = the K / N sizes involved."""
    P = dist.get_world_size(group)
    flags_needed = max(rows_total // 8, (rows_total // P // 128) * 4 * ((max(dims) + 63) // 64) * P) * 4
    return (rows_total % (128 * P) == 0 and all(int(d) % 8 == 0 for d in dims) and P <= 8
            and flags_needed <= _FLAG_BYTES and rows_total // 8 < 200000)   # per-unit flags stay below kAgDoneBase


# Experimental (QN_TP_AG_MODE=ce, round-2 validation pending): hand the all-gather to the COPY ENGINE.  Right after
# the producer of x_local, a side stream copies the shard into every peer's gather buffer with cudaMemcpyAsync and
# publishes one flag per destination; the AG->GEMM kernel then runs without push warps and only acquires that flag
# before its first remote tile.  No SM, register or L1 involvement in the transfer.

def _ag_mode() -> str:
    """Read per call (the micro-benchmarks switch modes at run time); must mirror fill_comm() in csrc/gemm.cu."""
    return {"c": "ce", "l": "lsu"}.get((os.environ.get("QN_TP_AG_MODE") or "pump")[:1], "pump")


_cudart = None


def _copy_engine_allgather(ws: "_Workspace", x_local: torch.Tensor, par: int, rows: int, K: int):
    global _cudart
    if _cudart is None:
        _cudart = ctypes.CDLL("libcudart.so")
    lib = native.require()
    if getattr(ws, "side", None) is None:
        ws.side = torch.cuda.Stream(device=x_local.device)
        ws.ag_done_base = int(lib.qn_ag_done_base())
    cur = torch.cuda.current_stream()
    ws.side.wait_stream(cur)
    nbytes = rows * K * 2
    off = ws.rank * nbytes
    with torch.cuda.stream(ws.side):
        stream = ctypes.c_void_p(ws.side.cuda_stream)
        for p in range(ws.world):
            dst = ws.gather[par].peer_ptrs[p] + off
            rc = _cudart.cudaMemcpyAsync(ctypes.c_void_p(dst), ctypes.c_void_p(x_local.data_ptr()),
                                         ctypes.c_size_t(nbytes), 3, stream)
            if rc != 0:
                raise RuntimeError(f"cudaMemcpyAsync (copy-engine all-gather) failed: {rc}")
            if p != ws.rank:
                native.check(lib.qn_write_flag(ws.ag_flags.peer_ptrs[p], ws.ag_done_base + ws.rank, ws.seq_ag, stream),
                             "qn_write_flag")
        ev = ws.side.record_event()
    x_local.record_stream(ws.side)
    return ev


def _ag_gemm(ws: _Workspace, x_local: torch.Tensor, b: torch.Tensor, b_mn: bool, N: int, K: int, epilogue: int,
             bias, z, out_full_x: bool, transient: bool = False):
    """returns (Y [M, N], X_full [M, K] or None).

    ``transient=True`` (backward): the kernel leaves the complete gathered operand in this rank's symmetric gather
    buffer (peers' rows arrive by NVLink, own rows are written by the same copy pumps) and that buffer is returned
    as a view — valid until the next-but-one all-gather launch of this workspace recycles it, i.e. long enough for
    the weight-gradient GEMM that immediately follows.  Otherwise (forward: the operand must survive until
    backward) the pumps assemble it in a private tensor while the GEMM runs — no separate clone of the buffer."""
    rows = x_local.shape[0]
    M = rows * ws.world
    ws.ensure(M * K * 2)
    ws.seq_ag += 1
    par = ws.seq_ag & 1
    y = torch.empty((M, N), dtype=torch.bfloat16, device=x_local.device)
    ce_event = _copy_engine_allgather(ws, x_local, par, rows, K) if _ag_mode() == "ce" else None
    x_priv = None
    if out_full_x and not transient and _ag_mode() == "pump":
        x_priv = torch.empty((M, K), dtype=torch.bfloat16, device=x_local.device)
    rc = native.require().qn_gemm_ag_bf16(
        x_local.data_ptr(), ws.ptrs(ws.gather[par]), ws.ptrs(ws.ag_flags), ws.rank, ws.world, ws.seq_ag,
        b.data_ptr(), y.data_ptr(), M, N, K, b.stride(0), N, int(b_mn), epilogue, native.ptr(bias), native.ptr(z),
        z.stride(0) if z is not None else 0, 0, 0, native.ptr(x_priv), native.stream())
    native.check(rc, f"qn_gemm_ag_bf16(M={M},N={N},K={K})")
    native.count()
    if x_priv is not None:
        return y, x_priv
    x_full = None
    if out_full_x:
        if ce_event is not None:               # the local rows reach the gather buffer on the side stream
            torch.cuda.current_stream().wait_event(ce_event)
        x_full = ws.gather[par].tensor[: M * K * 2].view(torch.bfloat16).view(M, K)
        if not transient:
            x_full = x_full.clone()
    return y, x_full


def _gemm_rs(ws: _Workspace, a: torch.Tensor, b: torch.Tensor, b_mn: bool, N: int, K: int, bias):
    """returns Y_local [M/world, N]"""
    M = a.shape[0]
    ws.ensure(M * ((N + 63) // 64) * 128)            # [world][M/world/32 * ceil(N/64)] slab images of 4 KB
    ws.seq_rs += 1
    par = ws.seq_rs & 1
    y = torch.empty((M // ws.world, N), dtype=torch.bfloat16, device=a.device)
    rc = native.require().qn_gemm_rs_bf16(
        a.data_ptr(), b.data_ptr(), y.data_ptr(), ws.ptrs(ws.stage[par]), ws.ptrs(ws.rs_flags), ws.rank, ws.world,
        ws.seq_rs, M, N, K, a.stride(0), b.stride(0), N, int(b_mn), native.ptr(bias), 0, 0, native.stream())
    native.check(rc, f"qn_gemm_rs_bf16(M={M},N={N},K={K})")
    native.count()
    return y


class _AllGatherLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, group, act):
        K, N = x.shape[-1], w.shape[0]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        ws = _ws(group, x.device)
        M = x2.shape[0] * ws.world
        z = torch.empty((M, N), dtype=torch.bfloat16, device=x.device) if act == "gelu" else None
        epi = EPI_BIAS_GELU if act == "gelu" else (EPI_BIAS if b is not None else EPI_NONE)
        need_x = ctx.needs_input_grad[1]
        y, x_full = _ag_gemm(ws, x2, w, False, N, K, epi, b, z, need_x)
        ctx.save_for_backward(x_full, w, z)
        ctx.group, ctx.has_bias, ctx.xshape = group, b is not None, x.shape
        ctx.b_param = b
        return y.view(x.shape[0] * ws.world, *x.shape[1:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x_full, w, z = ctx.saved_tensors
        N, K = w.shape
        dy2 = dy.reshape(-1, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        if z is not None:
            dy2 = dy2 * z                                  # z = gelu'(pre-activation)
        M = dy2.shape[0]
        ws = _ws(ctx.group, dy.device)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:                        # dX_local = RS(dY · W)
            dx = _gemm_rs(ws, dy2, w, True, K, N, None).view(ctx.xshape)
        if ctx.needs_input_grad[1]:
            dw = wgrad(w, dy2, x_full, N, K, M)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = bias_grad(ctx.b_param, dy2)
        return dx, dw, db, None, None


class _LinearReduceScatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, group):
        K, N = x.shape[-1], w.shape[0]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        ws = _ws(group, x.device)
        y = _gemm_rs(ws, x2, w, False, N, K, b)
        ctx.save_for_backward(x2, w)
        ctx.group, ctx.has_bias, ctx.xshape = group, b is not None, x.shape
        ctx.b_param = b
        return y.view(x.shape[0] // ws.world, *x.shape[1:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        N, K = w.shape
        dy2 = dy.reshape(-1, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        ws = _ws(ctx.group, dy.device)
        M = x2.shape[0]
        need_full = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        # dX = AG(dY_local) · W   (W [N, K] stored [K_red = N, N_out = K] -> MN-major B)
        dx, dy_full = _ag_gemm(ws, dy2, w, True, K, N, EPI_NONE, None, None, need_full, transient=True)
        dw = wgrad(w, dy_full, x2, N, K, M) if ctx.needs_input_grad[1] else None
        db = bias_grad(ctx.b_param, dy_full) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx.view(ctx.xshape), dw, db, None


class _SequenceParallelMLP(torch.autograd.Function):
    """Whole tensor-parallel MLP on token-sharded activations: 2 fused launches forward
    (AG->GEMM+bias+GELU, GEMM->RS+bias) and 2 backward (AG->GEMM ⊙ gelu', GEMM->RS) plus local wgrads."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, group):
        E, I = x.shape[-1], w1.shape[0]
        x2 = x.reshape(-1, E)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        ws = _ws(group, x.device)
        M = x2.shape[0] * ws.world
        g = torch.empty((M, I), dtype=torch.bfloat16, device=x.device)           # gelu'(pre-activation)
        a, x_full = _ag_gemm(ws, x2, w1, False, I, E, EPI_BIAS_GELU, b1, g, True)
        y = _gemm_rs(ws, a, w2, False, w2.shape[0], I, b2)
        ctx.save_for_backward(x_full, w1, w2, g, a)
        ctx.group, ctx.has_b1, ctx.has_b2, ctx.xshape = group, b1 is not None, b2 is not None, x.shape
        ctx.b1_param, ctx.b2_param = b1, b2
        return y.view(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x_full, w1, w2, g, a = ctx.saved_tensors
        I, E = w1.shape
        O = w2.shape[0]
        dy2 = dy.reshape(-1, O)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        ws = _ws(ctx.group, dy.device)
        M = x_full.shape[0]
        # dZ = (AG(dY_local) · W2) ⊙ gelu'(z)   — gelu' applied in the fused kernel's epilogue
        dz, dy_full = _ag_gemm(ws, dy2, w2, True, I, O, EPI_MUL_AUX, None, g, True, transient=True)
        dw2 = wgrad(w2, dy_full, a, O, I, M)
        db2 = bias_grad(ctx.b2_param, dy_full) if ctx.has_b2 else None
        dx = _gemm_rs(ws, dz, w1, True, E, I, None).view(ctx.xshape)             # dX_local = RS(dZ · W1)
        dw1 = wgrad(w1, dz, x_full, I, E, M)
        db1 = bias_grad(ctx.b1_param, dz) if ctx.has_b1 else None
        return dx, dw1, db1, dw2, db2, None


def sequence_parallel_mlp(x, w1, b1, w2, b2, group):
    return _SequenceParallelMLP.apply(x, w1, b1, w2, b2, group)


def allgather_linear(x, w, b, group, act: Optional[str] = None):
    return _AllGatherLinear.apply(x, w, b, group, act)


def linear_reduce_scatter(x, w, b, group):
    return _LinearReduceScatter.apply(x, w, b, group)
