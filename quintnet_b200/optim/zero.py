"""ZeRO stage helpers.  Stage 1 (optimizer-state sharding) is implemented by
:class:`~quintnet_b200.optim.distributed_adamw.DistributedAdamW`; this module
offers the convenience constructor used by trainers and strategies."""
from __future__ import annotations

from ..parallel.data_parallel import BucketConfig, DataParallel, DistributedConfig
from .distributed_adamw import DistributedAdamW


def wrap_zero1(model, dp_group, dp_rank: int, dp_size: int, capacity_mb: int = 64, **ddp_kw) -> DataParallel:
    """Wrap ``model`` in a DataParallel whose buckets are ZeRO-1 ready
    (flat parameters, deferred reduction)."""
    return DataParallel(model, DistributedConfig(rank=dp_rank, world_size=dp_size, process_group=dp_group),
                        BucketConfig(capacity_mb=capacity_mb, flatten_parameters=True),
                        reduce_mode="defer", **ddp_kw)


def zero1_optimizer(ddp: DataParallel, **kw) -> DistributedAdamW:
    return DistributedAdamW(ddp, **kw)
