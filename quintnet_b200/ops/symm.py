"""NVLink symmetric memory: allocation, rendezvous and the in-kernel collectives
built on it (``csrc/comm.cu``).

A *symmetric buffer* is an allocation that exists at the same size on every
rank of a process group, with every peer's copy mapped into this process'
address space (and, when the fabric supports NVLS, a multicast address that
fans out to / reduces over all copies inside the NVSwitch).  The sm_100a
kernels take the raw peer pointers and do the communication themselves with
ordinary ``ld``/``st`` (or ``multimem.*``) — no NCCL call on the data path.
(The reference has no counterpart: all of its collectives are ``torch.distributed`` NCCL calls on the default
stream — ``core/communication.py:416-424,515,576``, ``parallelism/data_parallel/backends/torch_backend.py:59,71``.)

Rendezvous back-ends, tried in order:
  1. ``torch.distributed._symmetric_memory`` (CUDA VMM + fabric/fd handles;
     also provides the multicast pointer and a signal pad);
  2. plain CUDA IPC (``cudaIpcGetMemHandle`` on a dedicated ``cudaMalloc``
     block, handles exchanged over the process group) — peer pointers only.
"""
from __future__ import annotations

import bisect
import ctypes
import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import native

_P, _I, _L, _F, _U32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_uint
native.declare("qn_allreduce", [_P, _P, _P, _I, _I, _L, _I, _F, _U32, _I, _P])
native.declare("qn_reduce_scatter_sqnorm", [_P, _P, _P, _I, _I, _P, _P, _P, _L, _I, _F, _U32, _I, _P])
native.declare("qn_allreduce_oneshot", [_P, _P, _P, _I, _I, _P, _P, _L, _I, _F, _P, _L, _I, _P])
native.declare("qn_adamw_allgather", [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _L, _I, _F, _F, _F, _F, _F, _F, _F,
                                      _U32, _I, _P])
native.declare("qn_ipc_get_handle", [_P, _P])
native.declare("qn_ipc_open_handle", [_P, ctypes.POINTER(ctypes.c_void_p)])
native.declare("qn_malloc", [ctypes.POINTER(ctypes.c_void_p), _L])
native.declare("qn_memset0", [_P, _L, _P])

SIGNAL_PAD_BYTES = 64 * 1024
COMM_BLOCKS = int(os.environ.get("QUINTNET_COMM_BLOCKS", "48"))      # <= 64 (signal-pad layout of csrc/comm.cu)
ONESHOT_MAX_BYTES = int(os.environ.get("QUINTNET_ONESHOT_MAX_BYTES", str(512 * 1024)))
_DT = {torch.float32: 0, torch.bfloat16: 1}


class _RawCudaTensorHolder:
    """Expose a raw device allocation to torch through ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3, "strides": None}


class SymmBuffer:
    """One symmetric allocation (``nbytes`` on every rank of ``group``)."""

    def __init__(self, tensor: torch.Tensor, group, peer_ptrs: List[int], sig_ptrs: List[int],
                 mc_ptr: int, backend: str, keepalive=None):
        self.tensor, self.group = tensor, group
        self.base = tensor.data_ptr()
        self.nbytes = tensor.numel() * tensor.element_size()
        self.peer_ptrs, self.sig_ptrs, self.mc_ptr = peer_ptrs, sig_ptrs, mc_ptr
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.backend = backend
        self.epoch = 2
        self._keepalive = keepalive

    def ptr_arrays(self, byte_offset: int = 0):
        arr = (ctypes.c_void_p * self.world)(*[p + byte_offset for p in self.peer_ptrs])
        sig = (ctypes.c_void_p * self.world)(*self.sig_ptrs)
        return arr, sig

    def next_epoch(self, n: int = 2) -> int:
        e = self.epoch
        self.epoch = (self.epoch + n) & 0x7FFFFFFF
        if self.epoch < 2:
            self.epoch = 2
        return e


_registry: List[SymmBuffer] = []
_sorted_bases = None            # (sorted base pointers, buffers in that order): lookup() is a bisect, not a scan
_pads: Dict[Tuple[int, str], SymmBuffer] = {}
_disabled_reason: Optional[str] = None
_pad_configured = False


def lookup(t: torch.Tensor) -> Optional[Tuple[SymmBuffer, int]]:
    """``(buffer, byte_offset)`` if ``t`` lives inside a symmetric allocation."""
    if not t.is_cuda:
        return None
    p = t.data_ptr()
    global _sorted_bases
    if _sorted_bases is None or len(_sorted_bases[0]) != len(_registry):     # rebuilt only when a buffer was added
        order = sorted(_registry, key=lambda b: b.base)
        _sorted_bases = ([b.base for b in order], order)
    i = bisect.bisect_right(_sorted_bases[0], p) - 1
    if i >= 0:
        b = _sorted_bases[1][i]
        if p < b.base + b.nbytes:
            return b, p - b.base
    return None


def _group_key(group) -> int:
    return id(group)


def _torch_symm_alloc(nbytes: int, device, group):
    import torch.distributed._symmetric_memory as sm
    global _pad_configured
    if not _pad_configured:
        _pad_configured = True
        try:
            if sm.get_signal_pad_size() < SIGNAL_PAD_BYTES:
                sm.set_signal_pad_size(SIGNAL_PAD_BYTES)
        except Exception:  # noqa: BLE001 — older torch / allocations already made: the default pad (9216 B) is enough
            pass
    t = sm.empty(nbytes, dtype=torch.uint8, device=device)
    try:
        h = sm.rendezvous(t, group=group)
    except TypeError:
        h = sm.rendezvous(t, group.group_name)
    world = dist.get_world_size(group)
    peer = [int(h.buffer_ptrs[i]) for i in range(world)]
    sig = [int(h.signal_pad_ptrs[i]) for i in range(world)]
    mc = 0
    try:
        mc = int(h.multicast_ptr or 0)
    except Exception:  # noqa: BLE001 — hardware/driver probe: any failure just means no NVLS multicast here
        mc = 0
    return t, peer, sig, mc, h


def _ipc_alloc(nbytes: int, device, group):
    lib = native.require()
    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def one(nb):
        out = ctypes.c_void_p()
        native.check(lib.qn_malloc(ctypes.byref(out), nb), "cudaMalloc")
        native.check(lib.qn_memset0(out.value, nb, native.stream()), "cudaMemset")
        hb = ctypes.create_string_buffer(64)
        native.check(lib.qn_ipc_get_handle(out.value, hb), "cudaIpcGetMemHandle")
        handles = [None] * world
        dist.all_gather_object(handles, hb.raw, group=group)
        ptrs = []
        for r in range(world):
            if r == rank:
                ptrs.append(out.value)
            else:
                o = ctypes.c_void_p()
                native.check(lib.qn_ipc_open_handle(handles[r], ctypes.byref(o)), "cudaIpcOpenMemHandle")
                ptrs.append(o.value)
        return out.value, ptrs

    torch.cuda.synchronize()
    base, peer = one(nbytes)
    sbase, sig = one(SIGNAL_PAD_BYTES)
    torch.cuda.synchronize()
    dist.barrier(group=group)
    holder = _RawCudaTensorHolder(base, nbytes)
    t = torch.as_tensor(holder, device=device)
    return t, peer, sig, 0, holder


def alloc_bytes(nbytes: int, device, group) -> SymmBuffer:
    global _disabled_reason
    nbytes = (nbytes + 4095) // 4096 * 4096
    mode = os.environ.get("QUINTNET_SYMM", "auto")
    errs = []
    if mode in ("auto", "torch"):
        try:
            t, peer, sig, mc, keep = _torch_symm_alloc(nbytes, device, group)
            if os.environ.get("QUINTNET_NVLS", "1") != "1":
                mc = 0
            b = SymmBuffer(t, group, peer, sig, mc, "torch_symm_mem", keep)
            _registry.append(b)
            return b
        except Exception as e:  # noqa: BLE001
            errs.append(f"torch symm_mem: {type(e).__name__}: {e}")
            if mode == "torch":
                raise
    try:
        t, peer, sig, mc, keep = _ipc_alloc(nbytes, device, group)
        b = SymmBuffer(t, group, peer, sig, mc, "cuda_ipc", keep)
        _registry.append(b)
        return b
    except Exception as e:  # noqa: BLE001
        errs.append(f"cuda ipc: {type(e).__name__}: {e}")
        _disabled_reason = "; ".join(errs)
        raise RuntimeError(f"symmetric memory unavailable: {_disabled_reason}")


def alloc(n: int, dtype: torch.dtype, device, group) -> torch.Tensor:
    """Zero-initialised symmetric 1-D tensor of ``n`` elements of ``dtype``."""
    es = torch.empty((), dtype=dtype).element_size()
    b = alloc_bytes(n * es, device, group)
    t = b.tensor[: n * es].view(dtype)
    t.zero_()
    return t


def available(group) -> bool:
    return (group is not None and dist.get_world_size(group) > 1 and torch.cuda.is_available()
            and native.available() and _disabled_reason is None)


# ------------------------------------------------------------------ collectives
def set_flag_watchdog(seconds: float, sm_clock_hz: float = 1.9e9) -> None:
    """Deadline of the in-kernel peer-flag waits (default ≈ 20 s).  A wait that outlives it traps, which turns
    a dead or diverged peer into a CUDA error on the survivors instead of a hung GPU."""
    lib = native.require()
    cycles = int(seconds * sm_clock_hz)
    for fn in ("qn_comm_set_flag_watchdog", "qn_gemm_set_flag_watchdog"):
        native.check(getattr(lib, fn)(cycles), fn)


def _fault_injected(kind: str, rank: int) -> bool:
    """Test hook: ``QUINTNET_FAULT=drop_collective:<rank>`` makes that rank silently skip its next in-kernel
    collective — the peers' flag waits must then hit the watchdog."""
    spec = os.environ.get("QUINTNET_FAULT", "")
    if not spec:
        return False
    what, _, who = spec.partition(":")
    return what == kind and (who == "" or int(who) == rank)


def all_reduce_(buf_off: Tuple[SymmBuffer, int], t: torch.Tensor, scale: float = 1.0,
                use_multicast: bool = True, blocks: Optional[int] = None) -> None:
    """In-place SUM all-reduce (× ``scale``) of symmetric tensor ``t`` by K1."""
    b, off = buf_off
    if _fault_injected("drop_collective", b.rank):
        return
    arr, sig = b.ptr_arrays(off)
    mc = (b.mc_ptr + off) if (use_multicast and b.mc_ptr) else None
    rc = native.require().qn_allreduce(arr, sig, mc, b.rank, b.world, t.numel(), _DT[t.dtype], float(scale),
                                       b.next_epoch(2), blocks or COMM_BLOCKS, native.stream())
    native.check(rc, "qn_allreduce")
    native.count()


def reduce_scatter_sqnorm(flat_grad: torch.Tensor, out_shard: torch.Tensor, scale: float,
                          mask: Optional[torch.Tensor], sq: torch.Tensor) -> None:
    b, off = lookup(flat_grad)
    arr, sig = b.ptr_arrays(off)
    mc = (b.mc_ptr + off) if b.mc_ptr else None           # NVLS: the switch forms the sum (multimem.ld_reduce)
    rc = native.require().qn_reduce_scatter_sqnorm(arr, sig, mc, b.rank, b.world, out_shard.data_ptr(),
                                                   native.ptr(mask), sq.data_ptr(), flat_grad.numel(),
                                                   _DT[flat_grad.dtype], float(scale), b.next_epoch(2),
                                                   COMM_BLOCKS, native.stream())
    native.check(rc, "qn_reduce_scatter_sqnorm")
    native.count()


def adamw_allgather(flat_param: torch.Tensor, st: dict, coef: Optional[torch.Tensor], *, step: int, lr: float,
                    beta1: float, beta2: float, eps: float, weight_decay: float) -> None:
    b, off = lookup(flat_param)
    arr, sig = b.ptr_arrays(off)
    if coef is not None and coef.dtype != torch.float32:
        coef = coef.float()
    mc = (b.mc_ptr + off) if b.mc_ptr else None           # NVLS: one multimem.st replicates the new parameters
    rc = native.require().qn_adamw_allgather(arr, sig, mc, b.rank, b.world, st["grad"].data_ptr(),
                                             st["master"].data_ptr(), st["m"].data_ptr(), st["v"].data_ptr(),
                                             native.ptr(coef), flat_param.numel(), _DT[flat_param.dtype], lr,
                                             beta1, beta2, eps, weight_decay, 1.0 - beta1 ** step,
                                             1.0 - beta2 ** step, b.next_epoch(2), COMM_BLOCKS, native.stream())
    native.check(rc, "qn_adamw_allgather")
    native.count()


# ------------------------------------------------------------------ small-message one-shot all-reduce
class _OneShotWorkspace:
    def __init__(self, group, device):
        self.stage = alloc_bytes(2 * ONESHOT_MAX_BYTES, device, group)
        self.stage.tensor.zero_()
        self.counters = torch.zeros(64, dtype=torch.int32, device=device)
        torch.cuda.synchronize()
        dist.barrier(group=group)


_oneshot: Dict[int, _OneShotWorkspace] = {}


def oneshot_ok(x: torch.Tensor, group) -> bool:
    # TP-sized groups only: at 8 ranks the two-shot NVLS kernel (24 us) and NCCL (38 us) both beat the one-shot
    # kernel's 7 peer reads per element (48 us; profiles/r2_allreduce_sweep_8gpu.txt)
    return (x.is_cuda and x.dtype in _DT and x.numel() * x.element_size() <= ONESHOT_MAX_BYTES
            and "allreduce" in native._enabled_set() and available(group)
            and dist.get_world_size(group) <= 4)


def all_reduce_small(x: torch.Tensor, group, scale: float = 1.0) -> torch.Tensor:
    """Out-of-place SUM all-reduce (x ``scale``) of a small tensor in ONE launch and one NVLink round trip: copy-in to a
    symmetric staging half, one cross-rank flag barrier, NVLS ``multimem.ld_reduce`` (or a load per peer) straight into
    the result.  Replaces the NCCL all-reduce of the latency-bound TP activations (reference
    ``core/communication.py:495-518``); CUDA-graph replayable (the launch counter lives on the device)."""
    key = _group_key(group)
    if key not in _oneshot:
        _oneshot[key] = _OneShotWorkspace(group, x.device)
    ws = _oneshot[key]
    b = ws.stage
    xc = x.contiguous()
    vn = 8 if x.dtype == torch.bfloat16 else 4
    n = xc.numel()
    npad = (n + vn - 1) // vn * vn
    if npad != n or xc.data_ptr() % 16:
        buf = torch.zeros(npad, dtype=x.dtype, device=x.device)
        buf[:n].copy_(xc.reshape(-1))
        src = buf
        out = torch.empty(npad, dtype=x.dtype, device=x.device)
    else:
        src = xc
        out = torch.empty_like(xc)
    if _fault_injected("drop_collective", b.rank):
        return out.reshape(-1)[:n].view(x.shape)
    arr, sig = b.ptr_arrays(0)
    blocks = max(1, min(16, (npad // vn + 511) // 512))
    rc = native.require().qn_allreduce_oneshot(arr, sig, b.mc_ptr or None, b.rank, b.world, src.data_ptr(), out.data_ptr(),
                                               npad, _DT[x.dtype], float(scale), ws.counters.data_ptr(),
                                               ONESHOT_MAX_BYTES, blocks, native.stream())
    native.check(rc, "qn_allreduce_oneshot")
    native.count()
    return out.reshape(-1)[:n].view(x.shape)
