"""Token + position embedding (reference ``utils/GPT2/gpt2_embeddings.py:16-103``);
replicated across TP ranks."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ... import ops


class GPT2Embedding(nn.Module):
    def __init__(self, vocab_size: int, max_position_embeddings: int, embed_dim: int,
                 wte_weights: Optional[torch.Tensor] = None, wpe_weights: Optional[torch.Tensor] = None,
                 dropout_prob: float = 0.1, device=None, dtype=None, use_wpe: bool = True):
        super().__init__()
        self.vocab_size, self.max_position_embeddings, self.embed_dim = vocab_size, max_position_embeddings, embed_dim
        self.wte = nn.Embedding(vocab_size, embed_dim, device=device, dtype=dtype)
        self.wpe = nn.Embedding(max_position_embeddings, embed_dim, device=device, dtype=dtype) \
            if use_wpe else None
        self.dropout_prob = dropout_prob
        self.dropout = nn.Dropout(dropout_prob)
        with torch.no_grad():
            if wte_weights is not None:
                self.wte.weight.copy_(wte_weights)
            else:
                nn.init.normal_(self.wte.weight, std=0.02)
            if self.wpe is not None:
                if wpe_weights is not None:
                    self.wpe.weight.copy_(wpe_weights)
                else:
                    nn.init.normal_(self.wpe.weight, std=0.02)

    def forward(self, input_ids: torch.Tensor, position_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self.wte(input_ids)
        if self.wpe is not None:
            T = input_ids.shape[1]
            if position_ids is None:
                pos = self.wpe.weight[:T].unsqueeze(0)
            else:
                pos = self.wpe(position_ids)
            x = x + pos
        return ops.dropout_add(x, None, self.dropout_prob, self.training)
