"""Single-process GPT-2 (the reference's ``utils/GPT2/gpt2_model.py`` is an
empty file).  ``GPT2Model`` is a :class:`GPT2Stage` that is both first and last
stage; with a TP group it is the tensor-parallel model without pipelining."""
from __future__ import annotations

from typing import Optional

import torch

from .gpt2_config import GPT2Config
from .gpt2_stage import GPT2Stage


class GPT2Model(GPT2Stage):
    def __init__(self, config: GPT2Config, device=None, dtype=None, tp_group=None, tp_rank: int = 0,
                 tp_size: int = 1, seed: int = 0):
        from ...core.distributed_loading import random_gpt2_state_dict
        device = device if device is not None else torch.device("cpu")
        sd = random_gpt2_state_dict(config, tp_rank=tp_rank, tp_size=tp_size, pp_rank=0, pp_size=1,
                                    device=device, seed=seed, dtype=dtype)
        tmp = GPT2Stage.from_sharded_state_dict(sd, config, 0, 1, tp_rank, tp_size, tp_group, None,
                                                device, dtype=dtype)
        super().__init__(tmp.embedding, tmp.blocks, tmp.ln_f, None, True, True, None, config=config,
                         tp_group=tp_group)

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, max_new_tokens: int = 32, eos_token_id: Optional[int] = None):
        """Greedy decoding (full re-forward per token, as the reference's
        ``utils/metrics.py:74`` helper does)."""
        self.eval()
        out = input_ids
        for _ in range(max_new_tokens):
            ctx = out[:, -self.config.n_positions:]
            logits = self.forward(ctx)
            nxt = logits[:, -1].float().argmax(-1, keepdim=True)
            out = torch.cat([out, nxt], dim=1)
            if eos_token_id is not None and bool((nxt == eos_token_id).all()):
                break
        return out
