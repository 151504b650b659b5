"""Tensor-parallel GPT-2 MLP (reference ``utils/GPT2/gpt2_mlp.py:51-162``):
column-parallel ``c_fc`` with the exact-erf GELU fused into the GEMM epilogue,
row-parallel ``c_proj``."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ... import ops
from ...core.communication import All_Reduce, Copy_To_Parallel
from ...parallel.tensor_parallel.layers import ColumnParallelLinear, RowParallelLinear


class GPT2MLP(nn.Module):
    def __init__(self, embed_dim=None, hidden_dim: Optional[int] = None, tp_rank: int = 0, tp_size: int = 1,
                 tp_group=None, device=None,
                 c_fc_weight: Optional[torch.Tensor] = None, c_fc_bias: Optional[torch.Tensor] = None,
                 c_proj_weight: Optional[torch.Tensor] = None, c_proj_bias: Optional[torch.Tensor] = None,
                 dropout_prob: float = 0.1, dtype=None, config=None):
        """Reference signature (``utils/GPT2/gpt2_mlp.py:69-82``): ``embed_dim, hidden_dim, tp_rank, tp_size, tp_group,
        device, c_fc_weight, c_fc_bias, c_proj_weight, c_proj_bias=None, dropout_prob=0.1``.  Alternatively pass
        ``config=GPT2Config`` (or a config as the first argument) to take the sizes and options from it."""
        super().__init__()
        if config is None and hasattr(embed_dim, "n_embd"):
            config, embed_dim = embed_dim, None
        if config is None:
            from .gpt2_config import GPT2Config
            config = GPT2Config(n_embd=int(embed_dim), n_inner=int(hidden_dim) if hidden_dim else None,
                                n_head=max(int(embed_dim) // 64, 1), resid_pdrop=float(dropout_prob))
        self.embed_dim = config.n_embd
        self.inner_dim = config.n_inner if config.n_inner is not None else 4 * config.n_embd
        if self.inner_dim % tp_size:
            raise ValueError(f"n_inner={self.inner_dim} not divisible by tp={tp_size}")
        self.inner_dim_local = self.inner_dim // tp_size
        self.hidden_dim, self.hidden_dim_local = self.inner_dim, self.inner_dim_local      # reference attribute names
        self.tp_rank, self.tp_size, self.tp_group, self.device = tp_rank, tp_size, tp_group, device
        sp = bool(getattr(config, "sequence_parallel", False)) and tp_size > 1
        fp8 = self.fp8 = bool(getattr(config, "fp8_linear", False))
        if c_fc_bias is None and c_fc_weight is not None:
            c_fc_bias = torch.zeros(self.inner_dim_local)
        self.c_fc = ColumnParallelLinear(device, tp_group, self.embed_dim, self.inner_dim_local,
                                         weight_slice=c_fc_weight, bias_slice=c_fc_bias,
                                         gather_output=False, sequence_parallel=sp,
                                         activation="gelu", dtype=dtype, fp8=fp8)
        if c_proj_bias is None and c_proj_weight is None:
            c_proj_bias = torch.zeros(self.embed_dim)
        self.c_proj = RowParallelLinear(device, tp_group, self.inner_dim_local, self.embed_dim,
                                        weight_slice=c_proj_weight, bias_slice=c_proj_bias,
                                        input_is_parallel=True, sequence_parallel=sp, dtype=dtype, fp8=fp8)
        self.resid_pdrop = config.resid_pdrop
        self.dropout = nn.Dropout(config.resid_pdrop)

    def forward(self, hidden_states: torch.Tensor, residual: Optional[torch.Tensor] = None,
                raw: bool = False) -> torch.Tensor:
        if not self.c_fc.sequence_parallel and not self.fp8:
            # whole MLP as ONE autograd op (GELU and GELU' fused into GEMM epilogues); the TP conjugate
            # collectives f / g wrap it exactly as they wrap the two separate layers
            x = Copy_To_Parallel.apply(hidden_states, self.tp_group)
            h = ops.mlp(x, self.c_fc.proj.weight, self.c_fc.proj.bias, self.c_proj.proj.weight, self.c_proj.bias)
            h = All_Reduce.apply(h, self.tp_group)
            return h if raw else ops.dropout_add(h, residual, self.resid_pdrop, self.training)
        from ...parallel.tensor_parallel.layers import _fused_tp_ok
        rows_total = (hidden_states.numel() // hidden_states.shape[-1]) * self.tp_size
        if self.c_fc.sequence_parallel and not self.fp8 and \
                _fused_tp_ok(hidden_states, self.tp_group, rows_total, self.embed_dim, self.inner_dim_local):
            from ...ops import tp_fused
            h = tp_fused.sequence_parallel_mlp(hidden_states, self.c_fc.proj.weight, self.c_fc.proj.bias,
                                               self.c_proj.proj.weight, self.c_proj.bias, self.tp_group)
            return h if raw else ops.dropout_add(h, residual, self.resid_pdrop, self.training)
        h = self.c_fc(hidden_states)          # bias + GELU fused
        h = self.c_proj(h)
        return h if raw else ops.dropout_add(h, residual, self.resid_pdrop, self.training)
