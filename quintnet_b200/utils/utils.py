"""Small single-process helpers with the reference's signatures (``utils/utils.py``: ``setup_device`` :37,
``count_parameters`` :46, ``save_checkpoint`` :58, ``load_checkpoint`` :79)."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.nn as nn


def setup_device() -> torch.device:
    """This process' device: ``cuda:LOCAL_RANK`` when CUDA is available (the reference returns the bare ``cuda``
    device, which under torchrun is the same device after ``set_device``), else CPU."""
    if torch.cuda.is_available():
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)) % max(torch.cuda.device_count(), 1))
    return torch.device("cpu")


def count_parameters(model: nn.Module) -> int:
    """Number of trainable parameters."""
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def save_checkpoint(model: nn.Module, optimizer, epoch: int, loss: float, path: str) -> None:
    """``{epoch, model_state_dict, optimizer_state_dict, loss}`` in one file."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({"epoch": epoch, "model_state_dict": model.state_dict(),
                "optimizer_state_dict": None if optimizer is None else optimizer.state_dict(), "loss": loss}, path)


def load_checkpoint(model: nn.Module, optimizer, path: str, map_location="cpu") -> Tuple[nn.Module, object, int, float]:
    """Returns ``(model, optimizer, epoch, loss)`` like the reference."""
    ck = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(ck["model_state_dict"])
    if optimizer is not None and ck.get("optimizer_state_dict") is not None:
        optimizer.load_state_dict(ck["optimizer_state_dict"])
    return model, optimizer, ck.get("epoch", 0), ck.get("loss", None)
