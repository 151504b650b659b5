"""Pre-LN GPT-2 block (reference ``utils/GPT2/gpt2_block.py:57-187``);
residual adds are fused with the dropout of the producing branch."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ... import ops
from .gpt2_attention import GPT2Attention
from .gpt2_mlp import GPT2MLP


class GPT2Block(nn.Module):
    def __init__(self, config, tp_rank: int, tp_size: int, tp_group, device,
                 ln_1_weight=None, ln_1_bias=None, ln_2_weight=None, ln_2_bias=None,
                 attn_c_attn_weight=None, attn_c_attn_bias=None, attn_c_proj_weight=None,
                 attn_c_proj_bias=None, mlp_c_fc_weight=None, mlp_c_fc_bias=None,
                 mlp_c_proj_weight=None, mlp_c_proj_bias=None, dtype=None):
        super().__init__()
        E, eps = config.n_embd, config.layer_norm_epsilon
        self.hidden_dim = E
        self.ln_1 = nn.LayerNorm(E, eps=eps, device=device, dtype=dtype)
        self.attn = GPT2Attention(config, tp_rank, tp_size, tp_group, device, attn_c_attn_weight,
                                  attn_c_attn_bias, attn_c_proj_weight, attn_c_proj_bias, dtype=dtype)
        self.ln_2 = nn.LayerNorm(E, eps=eps, device=device, dtype=dtype)
        self.mlp = GPT2MLP(config=config, tp_rank=tp_rank, tp_size=tp_size, tp_group=tp_group, device=device,
                           c_fc_weight=mlp_c_fc_weight, c_fc_bias=mlp_c_fc_bias, c_proj_weight=mlp_c_proj_weight,
                           c_proj_bias=mlp_c_proj_bias, dtype=dtype)
        with torch.no_grad():
            for ln, w, b in ((self.ln_1, ln_1_weight, ln_1_bias), (self.ln_2, ln_2_weight, ln_2_bias)):
                if w is not None:
                    ln.weight.copy_(w)
                if b is not None:
                    ln.bias.copy_(b)

    def forward_fused(self, pending: Optional[torch.Tensor], residual: torch.Tensor):
        """Fused residual chain: takes the previous branch output ``pending`` (not yet dropped / added) and
        the residual stream, returns this block's (pending, residual).  Each dropout + residual-add is fused
        into the LayerNorm that consumes it (one kernel forward, one backward)."""
        p = self.attn.resid_pdrop
        if pending is None:
            a, h = ops.layer_norm(residual, self.ln_1.weight, self.ln_1.bias, self.ln_1.eps), residual
        else:
            a, h = ops.dropout_add_layer_norm(pending, residual, self.ln_1.weight, self.ln_1.bias, self.ln_1.eps,
                                              p, self.training)
        o = self.attn(a, raw=True)
        m, h = ops.dropout_add_layer_norm(o, h, self.ln_2.weight, self.ln_2.bias, self.ln_2.eps, p, self.training)
        return self.mlp(m, raw=True), h

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        h = hidden_states
        a = ops.layer_norm(h, self.ln_1.weight, self.ln_1.bias, self.ln_1.eps)
        h = self.attn(a, residual=h)
        m = ops.layer_norm(h, self.ln_2.weight, self.ln_2.bias, self.ln_2.eps)
        return self.mlp(m, residual=h)
