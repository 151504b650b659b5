"""Process bootstrap.  The reference ships stubs (``core/distributed.py:16-39``);
these are real: backend by device, eager NCCL init with ``device_id``, loopback
rendezvous defaults so single-node jobs work in containers."""
from __future__ import annotations

import datetime
import os
from typing import Optional

import torch
import torch.distributed as dist


def setup_distributed(backend: Optional[str] = None, init_method: Optional[str] = None,
                      timeout_seconds: int = 1800, device_type: Optional[str] = None) -> torch.device:
    """Reference signature ``(backend, init_method, timeout_seconds)`` (``core/distributed.py:16-31``, a stub there)
    plus ``device_type``; ``init_method`` defaults to the torchrun environment."""
    if device_type is None:
        device_type = "cuda" if torch.cuda.is_available() else "cpu"
    backend = backend or ("nccl" if device_type == "cuda" else "gloo")
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    device = torch.device("cuda", local_rank) if device_type == "cuda" else torch.device("cpu")
    if device_type == "cuda":
        torch.cuda.set_device(device)
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {"device_id": device} if device_type == "cuda" else {}
        if init_method:
            kwargs["init_method"] = init_method
        dist.init_process_group(backend=backend,
                                timeout=datetime.timedelta(seconds=timeout_seconds), **kwargs)
    return device


def cleanup_distributed() -> None:
    if dist.is_initialized():
        try:
            dist.barrier()
        except Exception:
            pass
        dist.destroy_process_group()


def get_rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def is_main_process() -> bool:
    return get_rank() == 0
