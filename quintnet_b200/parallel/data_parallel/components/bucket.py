"""Persistent flat gradient bucket.

Unlike the reference's ``GradientBucket`` (``components/bucket.py:32-180``:
``torch.cat`` to flatten, copy back to unflatten, ready counter never reset),
the bucket owns ONE contiguous buffer for its lifetime and every parameter's
``.grad`` is a view into it: reduction needs zero copies and the buffer can live
in NVLink symmetric memory for the in-kernel collectives.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

ALIGN_ELEMS = 128  # keep every parameter 256B-aligned inside the flat buffer


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def grad_of(p: torch.Tensor) -> Optional[torch.Tensor]:
    """The gradient of ``p`` as the optimizer will see it.  With high-precision main-grad buckets (bf16 parameters,
    fp32 bucket) that is ``p.main_grad`` — the fp32 bucket view every micro-batch gradient is accumulated into —
    and ``p.grad`` is only a transient of the current backward; otherwise it is ``p.grad`` (itself a bucket view).
    Everything that reads gradients between backward and the optimizer step (global-norm clipping, the tied-weight
    and sequence-parallel gradient syncs, optimizers) goes through this accessor."""
    mg = getattr(p, "main_grad", None)
    return mg if mg is not None else p.grad


class GradientBucket:
    def __init__(self, bucket_id: int, parameters: List[nn.Parameter], grad_dtype: Optional[torch.dtype] = None,
                 pad_to: int = 1, flatten_parameters: bool = False, allocator=None):
        params = list(parameters)
        grad_dtype = grad_dtype or params[0].dtype
        self.bucket_id = bucket_id
        self.parameters: List[nn.Parameter] = list(params)
        self.device = params[0].device
        self.param_dtype = params[0].dtype
        self.grad_dtype = grad_dtype
        self.offsets: List[int] = []
        off = 0
        for p in self.parameters:
            self.offsets.append(off)
            off += _round_up(p.numel(), ALIGN_ELEMS)
        self.numel = _round_up(max(off, 1), max(pad_to, 1) * ALIGN_ELEMS)
        alloc = allocator or (lambda n, dtype, device: torch.zeros(n, dtype=dtype, device=device))
        self.flat_grad: torch.Tensor = alloc(self.numel, grad_dtype, self.device)
        self.flat_grad.zero_()
        self.flat_param: Optional[torch.Tensor] = None
        self.grad_views: List[torch.Tensor] = [
            self.flat_grad[o:o + p.numel()].view(p.shape) for o, p in zip(self.offsets, self.parameters)
        ]
        if flatten_parameters:
            self.flat_param = alloc(self.numel, self.param_dtype, self.device)
            self.flat_param.zero_()
            with torch.no_grad():
                for o, p in zip(self.offsets, self.parameters):
                    view = self.flat_param[o:o + p.numel()].view(p.shape)
                    view.copy_(p.data)
                    p.data = view
        self._ready = 0
        self.launched = False
        self.handle = None
        self.attach_grads()

    # -- bookkeeping -------------------------------------------------------
    @property
    def size_bytes(self) -> int:
        return self.numel * self.flat_grad.element_size()

    @property
    def size_mb(self) -> float:
        return self.size_bytes / (1024 * 1024)

    def attach_grads(self) -> None:
        """(Re)point every ``param.grad`` at its slice of the flat buffer."""
        for p, v in zip(self.parameters, self.grad_views):
            if p.requires_grad and (p.grad is None or p.grad.data_ptr() != v.data_ptr()):
                if self.grad_dtype == p.dtype:
                    p.grad = v
                else:
                    p.grad = None  # low-precision param, high-precision bucket: hooks accumulate into main_grad
            if self.grad_dtype != p.dtype:
                p.main_grad = v

    def absorb(self, p: nn.Parameter, idx: int) -> None:
        """Make sure the freshly accumulated grad of ``p`` is inside the bucket."""
        v = self.grad_views[idx]
        g = p.grad
        if g is None:
            return
        if g.data_ptr() != v.data_ptr():
            if self.grad_dtype == p.dtype:
                v.copy_(g)
                p.grad = v
            else:
                v.add_(g)          # main-grad accumulation in bucket precision
                p.grad = None

    def increment_gradient_count(self) -> bool:
        self._ready += 1
        return self.are_all_gradients_ready()

    def are_all_gradients_ready(self) -> bool:
        return self._ready >= len(self.parameters)

    def reset(self) -> None:
        self._ready = 0
        self.launched = False
        self.handle = None

    def zero(self) -> None:
        self.flat_grad.zero_()

    # reference-compatible helpers (views make them trivial)
    def flatten_gradients(self) -> torch.Tensor:
        return self.flat_grad

    def unflatten_gradients(self, flattened_tensor: torch.Tensor) -> None:
        if flattened_tensor.data_ptr() != self.flat_grad.data_ptr():
            self.flat_grad.copy_(flattened_tensor)

    @property
    def is_ready(self) -> bool:       # reference attribute: every gradient of the bucket has arrived
        return self.are_all_gradients_ready()

    # ---- reference API (``components/bucket.py:94-180``) ----------------------------------------------------
    def has_gradients(self) -> bool:
        """True if at least one parameter of the bucket currently holds a gradient."""
        return any(p.grad is not None for p in self.parameters if p.requires_grad)

    def register_hooks(self, hook_fn) -> None:
        """Call ``hook_fn(param)`` after each parameter's gradient has been accumulated (post-accumulate hooks: the
        reference's tensor hooks fire BEFORE ``.grad`` is written — SURVEY §8-D1b)."""
        if not hasattr(self, "_hook_handles"):
            self._hook_handles = []
        for p in self.parameters:
            if p.requires_grad:
                self._hook_handles.append(p.register_post_accumulate_grad_hook(hook_fn))

    def remove_hooks(self) -> None:
        for h in getattr(self, "_hook_handles", []):
            h.remove()
        self._hook_handles = []

    def __repr__(self) -> str:
        return (f"GradientBucket(id={self.bucket_id}, params={len(self.parameters)}, "
                f"{self.size_mb:.2f} MB, {self.grad_dtype})")
