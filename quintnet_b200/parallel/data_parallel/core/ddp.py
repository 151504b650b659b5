"""``DataParallel``: bucketed, overlapped gradient synchronisation.

Public surface follows the reference (``parallelism/data_parallel/core/ddp.py:49-179``:
``.model``, ``forward``, ``backward`` passthrough for pipeline stages,
``remove_hooks``, ``get_bucket_info``) but the engine is new:

* gradients live in persistent flat buckets (optionally NVLink symmetric
  memory); ``param.grad`` is a view, so reduction is zero-copy;
* post-accumulate-grad hooks count readiness per backward pass; a bucket is
  reduced asynchronously the moment its last gradient lands (all-reduce by the
  sm_100a P2P/NVLS kernel on a comm stream, or NCCL/gloo ``async_op``), and an
  autograd-engine callback waits for all buckets at the end of backward;
* ``no_sync()`` / ``require_backward_grad_sync`` give micro-batch accumulation
  semantics (only the last micro-batch's backward communicates), which the
  pipeline schedules use — the reference has no such notion;
* ``reduce_mode='defer'`` leaves reduction to a ZeRO-1 optimizer that fuses
  reduce-scatter + AdamW + all-gather.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch
import torch.nn as nn
from torch.autograd import Variable

from ..backends.base import DistributedBackend
from ..backends.torch_backend import TorchDistributedBackend
from ..components import BucketManager, GradientReducer, ParameterBroadcaster
from .config import BucketConfig, DistributedConfig, ReductionStrategy


class DataParallel(nn.Module):
    def __init__(self, model: nn.Module,
                 distributed_config: Optional[DistributedConfig] = None,
                 bucket_config: Optional[BucketConfig] = None,
                 backend: Optional[DistributedBackend] = None,
                 reduction_strategy: ReductionStrategy = ReductionStrategy.MEAN,
                 reduce_mode: str = "all_reduce"):
        super().__init__()
        if reduce_mode not in ("all_reduce", "defer"):
            raise ValueError(f"unknown reduce_mode {reduce_mode!r}")
        self.model = model
        self.distributed_config = distributed_config or DistributedConfig()
        self.bucket_config = bucket_config or BucketConfig()
        self.backend = backend or self._default_backend()
        self.reduction_strategy = reduction_strategy
        self.reduce_mode = reduce_mode
        self.require_backward_grad_sync = True
        self._callback_queued = False
        self.bucket_manager = BucketManager(self.bucket_config)
        self.gradient_reducer = GradientReducer(self.backend, self.distributed_config, reduction_strategy)
        self.parameter_broadcaster = ParameterBroadcaster(self.backend, self.distributed_config)
        self._setup()

    # ------------------------------------------------------------------ setup
    def _default_backend(self) -> DistributedBackend:
        p = next(self.model.parameters(), None)
        if p is not None and p.is_cuda and self.distributed_config.world_size > 1:
            from ....ops import native
            if "allreduce" in native._enabled_set() and native.available():
                from ..backends.nccl import NVLinkBackend
                return NVLinkBackend()
        return TorchDistributedBackend()

    def _allocator(self):
        """Buckets go to NVLink symmetric memory when the native backend is active."""
        from ..backends.nccl import NVLinkBackend
        if isinstance(self.backend, NVLinkBackend):
            from ....ops import symm
            group = self.distributed_config.process_group
            state = {"ok": True}

            def alloc(n, dtype, device):
                if state["ok"]:
                    try:
                        return symm.alloc(n, dtype, device, group)
                    except Exception as e:  # noqa: BLE001 — loud, not silent
                        import sys
                        state["ok"] = False
                        print(f"[quintnet_b200] WARNING: symmetric memory unavailable ({e}); "
                              f"DP buckets fall back to NCCL collectives", file=sys.stderr, flush=True)
                return torch.zeros(n, dtype=dtype, device=device)

            return alloc
        return None

    def _setup(self) -> None:
        # parameters must be identical across replicas before bucket views are built
        self.parameter_broadcaster.broadcast_parameters(self.model)
        self.buckets = self.bucket_manager.create_buckets(
            self.model, pad_to=max(self.distributed_config.world_size, 1), allocator=self._allocator())
        self.bucket_manager.register_gradient_hooks(self._on_grad)

    # ------------------------------------------------------------------ hooks
    def _on_grad(self, bucket, param, idx) -> None:
        bucket.absorb(param, idx)
        if not self.require_backward_grad_sync:
            return
        if not self._callback_queued:
            Variable._execution_engine.queue_callback(self._finalize_backward)
            self._callback_queued = True
        if bucket.increment_gradient_count() and self.reduce_mode == "all_reduce" and not bucket.launched:
            self._on_bucket_ready(bucket)

    def _on_bucket_ready(self, bucket) -> None:
        self.gradient_reducer.reduce_bucket(bucket)

    def _finalize_backward(self) -> None:
        self._callback_queued = False
        if self.reduce_mode == "all_reduce":
            for b in self.buckets.values():
                if not b.launched:          # bucket with unused parameters
                    self.gradient_reducer.reduce_bucket(b)
            for b in self.buckets.values():
                if b.handle is not None:
                    b.handle.wait()
        self.bucket_manager.reset_buckets()

    def finish_gradient_sync(self) -> None:
        """Explicit barrier for callers that drive backward manually."""
        if self._callback_queued:
            self._finalize_backward()

    @contextlib.contextmanager
    def no_sync(self):
        prev = self.require_backward_grad_sync
        self.require_backward_grad_sync = False
        try:
            yield
        finally:
            self.require_backward_grad_sync = prev

    # --------------------------------------------------------------- nn.Module
    def forward(self, *args, **kwargs):
        return self.model(*args, **kwargs)

    def backward(self, *args, **kwargs):
        """Pass-through to the wrapped pipeline stage's custom backward."""
        return self.model.backward(*args, **kwargs)

    def zero_grad(self, set_to_none: bool = False) -> None:  # keep the bucket views alive
        for b in self.buckets.values():
            b.zero()
            b.attach_grads()

    def remove_hooks(self) -> None:
        self.bucket_manager.remove_hooks()

    def get_bucket_info(self):
        return self.bucket_manager.get_bucket_info()

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            # expose e.g. get_tensor_shapes / local_module of the wrapped module
            return getattr(super().__getattr__("model"), name)
