"""Tensor-parallel causal self-attention (reference ``utils/GPT2/gpt2_attention.py:24-181``).

``c_attn`` is column-parallel (no gather), attention runs over this rank's
heads, ``c_proj`` is row-parallel.  Unlike the reference the fused QKV shard is
**head-aware**: rank ``r`` holds ``[Q_r | K_r | V_r]``, so TP and
single-GPU models compute the same function.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ... import ops
from ...parallel.tensor_parallel.layers import ColumnParallelLinear, RowParallelLinear


class GPT2Attention(nn.Module):
    def __init__(self, config, tp_rank: int, tp_size: int, tp_group, device,
                 c_attn_weight: Optional[torch.Tensor] = None, c_attn_bias: Optional[torch.Tensor] = None,
                 c_proj_weight: Optional[torch.Tensor] = None, c_proj_bias: Optional[torch.Tensor] = None,
                 dtype=None):
        super().__init__()
        self.embed_dim, self.num_heads = config.n_embd, config.n_head
        self.head_dim = self.embed_dim // self.num_heads
        self.scale = self.head_dim ** -0.5
        if self.num_heads % tp_size:
            raise ValueError(f"n_head={self.num_heads} not divisible by tp={tp_size}")
        self.tp_rank, self.tp_size, self.tp_group, self.device = tp_rank, tp_size, tp_group, device
        self.num_heads_local = self.num_heads // tp_size
        self.embed_dim_local = self.embed_dim // tp_size
        sp = bool(getattr(config, "sequence_parallel", False)) and tp_size > 1
        fp8 = bool(getattr(config, "fp8_linear", False))
        if c_attn_bias is None and c_attn_weight is not None:
            c_attn_bias = torch.zeros(3 * self.embed_dim_local)
        self.c_attn = ColumnParallelLinear(device, tp_group, self.embed_dim, 3 * self.embed_dim_local,
                                           weight_slice=c_attn_weight, bias_slice=c_attn_bias,
                                           gather_output=False, sequence_parallel=sp, dtype=dtype, fp8=fp8)
        if c_proj_bias is None and c_proj_weight is None:
            c_proj_bias = torch.zeros(self.embed_dim)
        self.c_proj = RowParallelLinear(device, tp_group, self.embed_dim_local, self.embed_dim,
                                        weight_slice=c_proj_weight, bias_slice=c_proj_bias,
                                        input_is_parallel=True, sequence_parallel=sp, dtype=dtype, fp8=fp8)
        self.attn_pdrop, self.resid_pdrop = config.attn_pdrop, config.resid_pdrop
        self.attn_dropout = nn.Dropout(config.attn_pdrop)
        self.resid_dropout = nn.Dropout(config.resid_pdrop)
        self.use_rope = bool(getattr(config, "use_rope", False))
        self.rope_base = float(getattr(config, "rope_base", 10000.0))
        self._rope_cache = None

    def _rope(self, T, device):
        c = self._rope_cache
        if c is None or c[0].shape[0] < T or c[0].device != device:
            self._rope_cache = ops.rope_tables(max(T, 64), self.head_dim, self.rope_base, device)
            c = self._rope_cache
        return c[0][:T], c[1][:T]

    def forward(self, hidden_states: torch.Tensor, residual: Optional[torch.Tensor] = None,
                raw: bool = False) -> torch.Tensor:
        """``raw=True`` returns the c_proj output WITHOUT dropout/residual (the block fuses those into
        the following LayerNorm)."""
        qkv = self.c_attn(hidden_states)                      # [B, T, 3*E/tp]
        B, T, _ = qkv.shape
        Hl, d = self.num_heads_local, self.head_dim
        if not self.use_rope and ops.attention_qkvpacked_ok(qkv, d):
            # native tcgen05 flash attention straight on the packed projection (no split/transpose copies)
            o = ops.attention_qkvpacked(qkv, Hl, causal=True, dropout_p=self.attn_pdrop, training=self.training)
            o = self.c_proj(o)
            return o if raw else ops.dropout_add(o, residual, self.resid_pdrop, self.training)
        q, k, v = qkv.view(B, T, 3, Hl, d).permute(2, 0, 3, 1, 4)   # each [B, Hl, T, d] (views)
        if self.use_rope:
            cos, sin = self._rope(T, qkv.device)
            q, k = ops.rope(q, cos, sin), ops.rope(k, cos, sin)
        # (RoPE / head_dim != 64 variants: the packed tcgen05 kernel does not apply; long sequences use the library
        # kernel here, short ones the native shared-memory kernel)
        o = ops.attention(q, k, v, causal=True, dropout_p=self.attn_pdrop, training=self.training, allow_library=True)
        o = o.transpose(1, 2).reshape(B, T, Hl * d)
        o = self.c_proj(o)
        return o if raw else ops.dropout_add(o, residual, self.resid_pdrop, self.training)
