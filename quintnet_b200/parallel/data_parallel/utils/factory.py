"""Factories (reference ``parallelism/data_parallel/utils/factory.py:12-37``)."""
from typing import Optional

import torch

from ..backends.local_backend import LocalBackend
from ..core.config import BucketConfig, DistributedConfig
from ..core.ddp import DataParallel


def create_local_ddp(model: torch.nn.Module, bucket_config: Optional[BucketConfig] = None) -> DataParallel:
    return DataParallel(model, DistributedConfig(), bucket_config or BucketConfig(), backend=LocalBackend())


def create_distributed_ddp(model: torch.nn.Module, rank: int, world_size: int, process_group=None,
                           bucket_config: Optional[BucketConfig] = None) -> DataParallel:
    return DataParallel(model, DistributedConfig(rank=rank, world_size=world_size,
                                                 process_group=process_group),
                        bucket_config or BucketConfig())
