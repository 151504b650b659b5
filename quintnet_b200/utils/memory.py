"""Memory helpers (reference ``utils/memory.py:9-30``)."""
import gc

import torch


def get_memory_usage(device=None) -> dict:
    if not torch.cuda.is_available():
        return {"allocated": 0.0, "reserved": 0.0, "max_allocated": 0.0}
    return {"allocated": torch.cuda.memory_allocated(device) / 2 ** 20,
            "reserved": torch.cuda.memory_reserved(device) / 2 ** 20,
            "max_allocated": torch.cuda.max_memory_allocated(device) / 2 ** 20}


def clear_cache() -> None:
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
