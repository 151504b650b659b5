"""Small TP helpers (reference ``parallelism/tensor_parallel/utils.py:25-103``);
rank/size are taken from the TP *group* when one is given (the reference
returns global values)."""
from __future__ import annotations

import torch.distributed as dist


def get_tensor_parallel_rank(group=None) -> int:
    if not dist.is_initialized():
        return 0
    return dist.get_rank(group) if group is not None else dist.get_rank()


def get_tensor_parallel_world_size(group=None) -> int:
    if not dist.is_initialized():
        return 1
    return dist.get_world_size(group) if group is not None else dist.get_world_size()


def is_tensor_parallel_available() -> bool:
    return dist.is_available() and dist.is_initialized()


def synchronize(group=None) -> None:
    if dist.is_initialized():
        dist.barrier(group=group)


def print_rank_0(*args, **kwargs) -> None:
    if not dist.is_initialized() or dist.get_rank() == 0:
        print(*args, **kwargs)


def ensure_divisibility(numerator: int, denominator: int) -> None:
    if numerator % denominator != 0:
        raise ValueError(f"{numerator} is not divisible by {denominator}")
