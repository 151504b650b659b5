"""Communication backend interface for the DP engine (reference
``parallelism/data_parallel/backends/base.py:30-91``), extended with the async
bucket operations the overlapped engine needs."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch


class Handle:
    """Completion handle; ``wait()`` is stream-ordered on CUDA."""

    def wait(self) -> None:  # pragma: no cover - trivial
        pass


class DistributedBackend(ABC):
    @abstractmethod
    def is_initialized(self) -> bool: ...

    @abstractmethod
    def get_world_size(self, group=None) -> int: ...

    @abstractmethod
    def broadcast_tensor(self, tensor: torch.Tensor, src: int, group=None) -> None: ...

    @abstractmethod
    def all_reduce_tensor(self, tensor: torch.Tensor, op: str = "sum", group=None) -> None: ...

    def all_reduce_async(self, tensor: torch.Tensor, scale: float, group=None) -> Handle:
        """SUM all-reduce then multiply by ``scale``; may run on a side stream."""
        self.all_reduce_tensor(tensor, "sum", group)
        if scale != 1.0:
            tensor.mul_(scale)
        return Handle()

    def reduce_scatter_async(self, out: torch.Tensor, inp: torch.Tensor, scale: float,
                             group=None) -> Handle:
        raise NotImplementedError
