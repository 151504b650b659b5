"""DP configuration types (reference ``parallelism/data_parallel/core/config.py:36-84``).
Bucketing is always on here — the reference silently disables gradient
synchronisation unless ``gradient_as_bucket_view=True``."""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Optional

import torch
import torch.distributed as dist


class ReductionStrategy(Enum):
    SUM = "sum"
    MEAN = "mean"


@dataclass
class BucketConfig:
    capacity_mb: int = 25
    # grads are ALWAYS views into the flat bucket buffer in this implementation;
    # the flag is kept for API compatibility
    gradient_as_bucket_view: bool = True
    # parameters that receive no gradient in a step are ALWAYS tolerated: their bucket slots stay zero and the
    # end-of-backward callback reduces buckets whose ready-count never completed (API-compatibility flag)
    find_unused_parameters: bool = False
    # also make parameters views of a flat per-bucket buffer (needed by ZeRO-1)
    flatten_parameters: bool = False
    # dtype of the gradient buckets (None = parameter dtype)
    grad_dtype: Optional[torch.dtype] = None


@dataclass
class DistributedConfig:
    rank: int = 0
    world_size: int = 1
    process_group: Optional[dist.ProcessGroup] = None
    broadcast_buffers: bool = True
