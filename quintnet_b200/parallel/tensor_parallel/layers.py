"""Megatron-style tensor-parallel layers.

Constructor signatures, attribute names (``.proj``, ``.bias``) and state-dict
keys follow the reference (``parallelism/tensor_parallel/layers.py:42-297``) so
shards and checkpoints interchange.  What is new:

* mathematically complete backward: a column-parallel layer with a replicated
  input applies the conjugate *f* operator (identity fwd / all-reduce bwd)
  that the reference omits;
* ``sequence_parallel=True`` selects the token-sharded formulation —
  all-gather→GEMM for column, GEMM→reduce-scatter for row — which on B200 is
  served by ONE fused sm_100a kernel per layer (tcgen05 tiles + NVLink peer
  stores, :mod:`quintnet_b200.ops.tp_fused`), and by the unfused
  ``All_Gather``/``ReduceScatter`` autograd ops on CPU/gloo;
* the row-parallel bias (kept on tp-rank 0 only, as in the reference shards) is
  added to rank 0's *partial* product before the reduction, so every rank ends
  up with the same output (the reference adds it after the all-reduce on rank 0
  only, leaving the other ranks without a bias).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ... import ops
from ...core.communication import (All_Gather, All_Reduce, Copy_To_Parallel, ReduceScatter,
                                   Scatter_To_Parallel, group_rank, group_size)


def _fused_tp_ok(x: torch.Tensor, group, rows_total: int, *dims) -> bool:
    """True when the fused sm_100a GEMM+collective kernel can serve this call."""
    if not x.is_cuda or x.dtype != torch.bfloat16 or group_size(group) == 1:
        return False
    from ...ops import native
    if "tp_fused" not in native._enabled_set() or not native.available():
        return False
    from ...ops import tp_fused
    return tp_fused.available(group) and tp_fused.can_fuse(rows_total, group, *dims)


class ColumnParallelLinear(nn.Module):
    """Y_local = X · W_localᵀ + b_local, W split along the output dim."""

    def __init__(self, local_device, tp_group, in_features: int, out_features_per_rank: int,
                 weight_slice: Optional[torch.Tensor] = None,
                 bias_slice: Optional[torch.Tensor] = None,
                 gather_output: bool = True, sync_gradients: bool = True,
                 gather_mode: str = "slice", sequence_parallel: bool = False,
                 activation: Optional[str] = None, dtype: Optional[torch.dtype] = None, fp8: bool = False):
        super().__init__()
        self.device, self.tp_group = local_device, tp_group
        self.fp8 = fp8            # forward GEMM on row-wise-scaled e4m3 operands (SURVEY 2.3 "FP8 option")
        self.in_features, self.out_features_per_rank = in_features, out_features_per_rank
        self.gather_output, self.sync_gradients = gather_output, sync_gradients
        self.gather_mode, self.sequence_parallel = gather_mode, sequence_parallel
        self.activation = activation
        has_bias = bias_slice is not None or weight_slice is None
        self.proj = nn.Linear(in_features, out_features_per_rank, bias=has_bias,
                              device=local_device, dtype=dtype)
        with torch.no_grad():
            if weight_slice is not None:
                self.proj.weight.copy_(weight_slice)
            if bias_slice is not None:
                self.proj.bias.copy_(bias_slice)
        self.proj.weight._qn_tp_sharded = True
        if self.proj.bias is not None:
            self.proj.bias._qn_tp_sharded = True

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        w, b = self.proj.weight, self.proj.bias
        if self.sequence_parallel and group_size(self.tp_group) > 1:
            # x is token-sharded along dim 0: all-gather -> GEMM
            rows_total = (x.numel() // x.shape[-1]) * group_size(self.tp_group)
            if not self.fp8 and _fused_tp_ok(x, self.tp_group, rows_total, w.shape[0], w.shape[1]) and \
                    (b is None or b.dtype == torch.bfloat16):
                from ...ops import tp_fused
                return tp_fused.allgather_linear(x, w, b, self.tp_group, self.activation)
            xg = All_Gather.apply(x, self.tp_group, "reduce_scatter", 0)
            return ops.linear(xg, w, b, self.activation, fp8=self.fp8)
        if self.sync_gradients:
            x = Copy_To_Parallel.apply(x, self.tp_group)
        y = ops.linear(x, w, b, self.activation, fp8=self.fp8)
        if self.gather_output:
            y = All_Gather.apply(y, self.tp_group, self.gather_mode, -1)
        return y

    def extra_repr(self) -> str:
        return (f"in={self.in_features}, out_per_rank={self.out_features_per_rank}, "
                f"gather_output={self.gather_output}, sp={self.sequence_parallel}")


class RowParallelLinear(nn.Module):
    """Y = Σ_ranks X_local · W_localᵀ (+ b), W split along the input dim."""

    def __init__(self, local_device, tp_group, in_features_per_rank: int, out_features: int,
                 weight_slice: Optional[torch.Tensor] = None,
                 bias_slice: Optional[torch.Tensor] = None,
                 input_is_parallel: bool = True, sequence_parallel: bool = False,
                 dtype: Optional[torch.dtype] = None, fp8: bool = False):
        super().__init__()
        self.device, self.tp_group = local_device, tp_group
        self.fp8 = fp8
        self.in_features_per_rank, self.out_features = in_features_per_rank, out_features
        self.input_is_parallel, self.sequence_parallel = input_is_parallel, sequence_parallel
        self.proj = nn.Linear(in_features_per_rank, out_features, bias=False,
                              device=local_device, dtype=dtype)
        with torch.no_grad():
            if weight_slice is not None:
                self.proj.weight.copy_(weight_slice)
        # bias lives on tp-rank 0 only (reference shard format, layers.py:176-181)
        if bias_slice is not None and group_rank(tp_group) == 0:
            self.bias = nn.Parameter(bias_slice.detach().clone().to(device=local_device,
                                                                     dtype=dtype or bias_slice.dtype))
            self.bias._qn_tp_unique = True       # exists on tp-rank 0 only
        else:
            self.register_parameter("bias", None)
        self.proj.weight._qn_tp_sharded = True

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n = group_size(self.tp_group)
        if not self.input_is_parallel and n > 1:
            x = Scatter_To_Parallel.apply(x, self.tp_group, -1)
        w = self.proj.weight
        if self.sequence_parallel and n > 1:
            rows_total = x.numel() // x.shape[-1]
            if not self.fp8 and _fused_tp_ok(x, self.tp_group, rows_total, w.shape[0], w.shape[1]) and \
                    (self.bias is None or self.bias.dtype == torch.bfloat16):
                from ...ops import tp_fused
                return tp_fused.linear_reduce_scatter(x, w, self.bias, self.tp_group)
            y = ops.linear(x, w, self.bias, fp8=self.fp8)          # bias joins rank 0's partial sum
            return ReduceScatter.apply(y, self.tp_group, 0)
        y = ops.linear(x, w, self.bias, fp8=self.fp8)
        return All_Reduce.apply(y, self.tp_group)

    def extra_repr(self) -> str:
        return (f"in_per_rank={self.in_features_per_rank}, out={self.out_features}, "
                f"input_is_parallel={self.input_is_parallel}, sp={self.sequence_parallel}")


class VocabParallelEmbedding(nn.Module):
    """Embedding table split along the vocabulary (reference layers.py:224-297)."""

    def __init__(self, local_device, tp_group, num_embeddings: int, embedding_dim: int,
                 vocab_start_idx: int, vocab_end_idx: int,
                 weight_slice: Optional[torch.Tensor] = None, dtype: Optional[torch.dtype] = None):
        super().__init__()
        self.device, self.tp_group = local_device, tp_group
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.vocab_start_idx, self.vocab_end_idx = vocab_start_idx, vocab_end_idx
        n_local = vocab_end_idx - vocab_start_idx
        # the table lives in ``self.embedding`` (state-dict key ``embedding.weight``) as in the reference
        self.embedding = nn.Embedding(n_local, embedding_dim, device=local_device, dtype=dtype)
        with torch.no_grad():
            if weight_slice is not None:
                self.embedding.weight.copy_(weight_slice)
            else:
                nn.init.normal_(self.embedding.weight, std=0.02)
        self.embedding.weight._qn_tp_sharded = True

    @property
    def weight(self) -> nn.Parameter:
        return self.embedding.weight

    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        ids = input_ids
        in_range = (ids >= self.vocab_start_idx) & (ids < self.vocab_end_idx)
        local = (ids - self.vocab_start_idx).clamp(0, self.weight.shape[0] - 1)
        out = torch.nn.functional.embedding(local, self.weight)
        out = out * in_range.unsqueeze(-1).to(out.dtype)
        return All_Reduce.apply(out, self.tp_group)
