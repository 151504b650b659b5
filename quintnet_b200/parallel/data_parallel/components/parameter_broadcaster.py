"""Initial parameter (and buffer) broadcast from the first rank of the DP group
— one flat broadcast per dtype instead of one per parameter (reference
``components/parameter_broadcaster.py:50-77``), and buffers are honoured."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist
import torch.nn as nn

from ..backends.base import DistributedBackend
from ..core.config import DistributedConfig


class ParameterBroadcaster:
    def __init__(self, backend: DistributedBackend, distributed_config: DistributedConfig):
        self.backend, self.config = backend, distributed_config
        self.distributed_config = distributed_config

    def _src(self) -> int:
        g = self.config.process_group
        if g is not None and dist.is_initialized():
            return dist.get_global_rank(g, 0)
        return 0

    def _broadcast_flat(self, tensors: List[torch.Tensor]) -> None:
        by_key = {}
        for t in tensors:
            by_key.setdefault((t.device, t.dtype), []).append(t)
        for (_, _), ts in by_key.items():
            flat = torch.cat([t.detach().reshape(-1) for t in ts])
            self.backend.broadcast_tensor(flat, self._src(), self.config.process_group)
            off = 0
            with torch.no_grad():
                for t in ts:
                    t.copy_(flat[off:off + t.numel()].view_as(t))
                    off += t.numel()

    def broadcast_parameters(self, model: nn.Module) -> None:
        if self.backend.get_world_size(self.config.process_group) <= 1:
            return
        self._broadcast_flat([p.data for p in model.parameters()])
        if self.config.broadcast_buffers:
            bufs = [b for b in model.buffers() if b is not None and b.numel() > 0]
            if bufs:
                self._broadcast_flat(bufs)
