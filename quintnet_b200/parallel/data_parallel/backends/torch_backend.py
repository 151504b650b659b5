"""torch.distributed backend (NCCL on CUDA, gloo on CPU) with async bucket ops.

Reference: ``parallelism/data_parallel/backends/torch_backend.py:31-85`` (synchronous ``broadcast_tensor`` :59 /
``all_reduce_tensor`` :71 behind the ``DistributedBackend`` ABC of ``backends/base.py:30-91``).  Here the bucket
reduction is asynchronous (``async_op=True`` handle, scale applied on completion) so it overlaps with backward;
on B200 the hot path is :mod:`.nccl` (in-kernel NVLink all-reduce) and this class is the gloo / fallback path."""
from __future__ import annotations

import torch.distributed as dist

from .base import DistributedBackend, Handle


class _WorkHandle(Handle):
    def __init__(self, work, tensor, scale):
        self.work, self.tensor, self.scale = work, tensor, scale

    def wait(self) -> None:
        if self.work is not None:
            self.work.wait()
            self.work = None
            if self.scale != 1.0:
                self.tensor.mul_(self.scale)


def _scaled(out, scale):
    out.mul_(scale)
    return Handle()


class TorchDistributedBackend(DistributedBackend):
    def is_initialized(self) -> bool:
        return dist.is_available() and dist.is_initialized()

    def get_world_size(self, group=None) -> int:
        return dist.get_world_size(group) if self.is_initialized() else 1

    def broadcast_tensor(self, tensor, src, group=None) -> None:
        if self.is_initialized() and self.get_world_size(group) > 1:
            dist.broadcast(tensor, src=src, group=group)

    def all_reduce_tensor(self, tensor, op="sum", group=None) -> None:
        if self.is_initialized() and self.get_world_size(group) > 1:
            red = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op]
            dist.all_reduce(tensor, op=red, group=group)

    def all_reduce_async(self, tensor, scale, group=None) -> Handle:
        if not self.is_initialized() or self.get_world_size(group) == 1:
            if scale != 1.0:
                tensor.mul_(scale)
            return Handle()
        work = dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group, async_op=True)
        return _WorkHandle(work, tensor, scale)

    def reduce_scatter_async(self, out, inp, scale, group=None) -> Handle:
        if not self.is_initialized() or self.get_world_size(group) == 1:
            out.copy_(inp.view_as(out))
            if scale != 1.0:
                out.mul_(scale)
            return Handle()
        if inp.dtype != out.dtype:
            inp = inp.to(out.dtype)
        from ....core.communication import _reduce_scatter_tensor
        _reduce_scatter_tensor(out, inp, group)
        return _WorkHandle(None, out, 1.0) if scale == 1.0 else _scaled(out, scale)
