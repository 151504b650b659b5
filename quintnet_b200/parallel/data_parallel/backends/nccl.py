"""Native NVLink backend: bucket all-reduce by the hand-written sm_100a P2P /
NVLS kernel over symmetric memory (K1), launched on a communication stream so
it overlaps with the rest of backward.  (The reference's ``backends/nccl.py``
is an empty stub.)"""
from __future__ import annotations

import torch

from .base import Handle
from .torch_backend import TorchDistributedBackend


class _EventHandle(Handle):
    def __init__(self, event):
        self.event = event

    def wait(self) -> None:
        if self.event is not None:
            torch.cuda.current_stream().wait_event(self.event)
            self.event = None


class NVLinkBackend(TorchDistributedBackend):
    """Buckets must live in symmetric memory (see ``DataParallel``: buckets are
    allocated through :func:`quintnet_b200.ops.symm.alloc_like` when this
    backend is active)."""

    def __init__(self):
        self._comm_stream = None

    @property
    def comm_stream(self):
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(priority=-1)
        return self._comm_stream

    def all_reduce_async(self, tensor, scale, group=None) -> Handle:
        from ....ops import symm
        buf = symm.lookup(tensor)
        if buf is None or self.get_world_size(group) == 1:
            return super().all_reduce_async(tensor, scale, group)
        if self._tp_spin_kernels_active():
            # The in-kernel all-reduce spins on peer flags while it holds its SMs.  The fused TP GEMM+collective
            # kernels spin on peer flags too: co-resident on a dp x tp grid, each kind can keep the other's peer
            # kernel from getting the SMs it needs and close a cross-GPU wait cycle.  With fused TP active the
            # reduction therefore runs in stream order on the compute stream (no side-stream overlap).
            symm.all_reduce_(buf, tensor, scale)
            return _EventHandle(None)
        ready = torch.cuda.Event()
        ready.record()
        s = self.comm_stream
        s.wait_event(ready)
        with torch.cuda.stream(s):
            symm.all_reduce_(buf, tensor, scale)
            done = torch.cuda.Event()
            done.record(s)
        return _EventHandle(done)

    @staticmethod
    def _tp_spin_kernels_active() -> bool:
        import sys
        m = sys.modules.get("quintnet_b200.ops.tp_fused")
        return bool(m is not None and getattr(m, "_workspaces", None))
