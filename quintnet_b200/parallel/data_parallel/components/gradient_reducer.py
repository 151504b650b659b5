"""Bucket reduction (reference ``components/gradient_reducer.py:64-105`` was
synchronous, copy-heavy and swallowed exceptions; this one is async, zero-copy
and lets errors propagate)."""
from __future__ import annotations

from ..backends.base import DistributedBackend, Handle
from ..core.config import DistributedConfig, ReductionStrategy
from .bucket import GradientBucket


class GradientReducer:
    def __init__(self, backend: DistributedBackend, distributed_config: DistributedConfig,
                 reduction_strategy: ReductionStrategy = ReductionStrategy.MEAN):
        self.backend, self.config, self.reduction_strategy = backend, distributed_config, reduction_strategy
        self.distributed_config = distributed_config

    @property
    def scale(self) -> float:
        if self.reduction_strategy == ReductionStrategy.MEAN:
            return 1.0 / max(self.config.world_size, 1)
        return 1.0

    def reduce_bucket(self, bucket: GradientBucket) -> Handle:
        """Launch the (possibly asynchronous) reduction of one bucket."""
        h = self.backend.all_reduce_async(bucket.flat_grad, self.scale, self.config.process_group)
        bucket.handle, bucket.launched = h, True
        return h
