"""Small single-process helpers (reference ``utils/utils.py``: ``setup_device`` :37, ``count_parameters`` :46,
``save_checkpoint`` :58, ``load_checkpoint`` :79)."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .checkpoint import load_checkpoint, save_checkpoint  # noqa: F401  (same module path as in the reference)


def setup_device() -> torch.device:
    """This process' device: ``cuda:LOCAL_RANK`` when CUDA is available (the reference returns the bare ``cuda``
    device, which under torchrun is the same device after ``set_device``), else CPU."""
    if torch.cuda.is_available():
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)) % max(torch.cuda.device_count(), 1))
    return torch.device("cpu")


def count_parameters(model: nn.Module) -> int:
    """Number of trainable parameters."""
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
