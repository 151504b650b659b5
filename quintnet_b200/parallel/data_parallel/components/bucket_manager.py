"""Greedy reverse-order bucket assignment + post-accumulate-grad hooks
(reference ``components/bucket_manager.py:41-183`` used pre-accumulation tensor
hooks, which see the gradient *before* it is written — SURVEY §8-D1b)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn

from ..core.config import BucketConfig
from .bucket import GradientBucket


class BucketManager:
    def __init__(self, config: Optional[BucketConfig] = None):
        self.config = config or BucketConfig()
        self.buckets: Dict[int, GradientBucket] = {}
        self.param_to_bucket: Dict[nn.Parameter, tuple] = {}
        self._hooks = []

    def create_buckets(self, model: nn.Module, pad_to: int = 1, allocator=None) -> Dict[int, GradientBucket]:
        cap = int(self.config.capacity_mb * 1024 * 1024)
        params = [p for p in model.parameters() if p.requires_grad]
        seen, uniq = set(), []
        for p in reversed(params):  # grads arrive roughly in reverse registration order
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        groups: List[List[nn.Parameter]] = []
        cur: List[nn.Parameter] = []
        cur_bytes = 0
        cur_key = None
        for p in uniq:
            gd = self.config.grad_dtype or p.dtype
            key = (p.device, p.dtype, gd)
            nbytes = p.numel() * torch.empty((), dtype=gd).element_size()
            if cur and (key != cur_key or cur_bytes + nbytes > cap):
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_key = key
            cur_bytes += nbytes
        if cur:
            groups.append(cur)
        self.buckets = {}
        self.param_to_bucket = {}
        for i, g in enumerate(groups):
            gd = self.config.grad_dtype or g[0].dtype
            b = GradientBucket(i, g, gd, pad_to=pad_to,
                               flatten_parameters=self.config.flatten_parameters, allocator=allocator)
            self.buckets[i] = b
            for j, p in enumerate(g):
                self.param_to_bucket[p] = (b, j)
        return self.buckets

    def register_gradient_hooks(self, on_grad: Callable[[GradientBucket, nn.Parameter, int], None]) -> None:
        self.remove_hooks()
        for p, (b, j) in self.param_to_bucket.items():
            def hook(param, _b=b, _j=j):
                on_grad(_b, param, _j)
            self._hooks.append(p.register_post_accumulate_grad_hook(hook))

    def remove_hooks(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

    # ---- reference API names (``components/bucket_manager.py:117-170``) ------------------------------------
    def register_hooks(self, bucket_ready_callback: Callable[[int], None]) -> None:
        """``bucket_ready_callback(bucket_id)`` fires when every gradient of a bucket has been accumulated."""
        def on_grad(bucket, param, idx):
            bucket.absorb(param, idx)
            if bucket.increment_gradient_count():
                bucket_ready_callback(bucket.bucket_id)
        self.register_gradient_hooks(on_grad)

    def remove_all_hooks(self) -> None:
        self.remove_hooks()

    def reset_buckets(self) -> None:
        for b in self.buckets.values():
            b.reset()

    def get_bucket_info(self):
        return [{"id": b.bucket_id, "num_params": len(b.parameters), "size_mb": b.size_mb,
                 "dtype": str(b.grad_dtype)} for b in self.buckets.values()]
