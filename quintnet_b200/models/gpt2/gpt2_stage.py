"""One pipeline stage of GPT-2 (reference ``utils/GPT2/gpt2_stage.py:32-276``).

State-dict schema is the reference's (``embedding.wte.weight``,
``blocks.{i}.attn.c_attn.proj.weight`` …, ``ln_f.*``, ``lm_head_weight``).
New behaviour:

* ``forward(x, position_ids=None, labels=None)``: with ``labels`` the last
  stage returns ``(loss_sum, n_valid)`` from the fused lm_head+cross-entropy op
  — the ``[B, T, V]`` fp32 logits (3.3 GB at 32×512×50257) are never built
  (SURVEY §8-D14); labels are shifted for next-token prediction (D5);
* tied wte/lm_head gradients are SUM-reduced between the first and the last
  stage only, *before* the optimizer step (D6);
* token-sharded (sequence-parallel) TP: the residual stream carries
  ``B/tp`` sequences per TP rank, the lm_head+CE is split by tokens across the
  TP group instead of being replicated, and replicated parameters get their
  gradients summed over TP (:meth:`sync_sequence_parallel_grads`).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from ... import ops
from ...core.communication import All_Gather, group_rank, group_size
from ...parallel.data_parallel.components.bucket import grad_of
from .gpt2_block import GPT2Block
from .gpt2_embeddings import GPT2Embedding


class GPT2Stage(nn.Module):
    supports_fused_loss = True

    def __init__(self, embedding: Optional[GPT2Embedding], blocks: nn.ModuleList,
                 ln_f: Optional[nn.LayerNorm], lm_head_weight: Optional[nn.Parameter],
                 is_first_stage: bool, is_last_stage: bool, pp_group=None, config=None,
                 tp_group=None, tied_group=None, tie_in_place: bool = True):
        super().__init__()
        self.embedding = embedding
        self.blocks = blocks
        self.ln_f = ln_f
        self.is_first_stage, self.is_last_stage = is_first_stage, is_last_stage
        self.pp_group, self.tp_group = pp_group, tp_group
        self.tied_group = tied_group if tied_group is not None else pp_group
        self.config = config
        self.shift_labels = True
        self.sequence_parallel = bool(getattr(config, "sequence_parallel", False)) and group_size(tp_group) > 1
        if is_first_stage and is_last_stage and embedding is not None and tie_in_place:
            # whole model on one stage: a real tie, no gradient sync needed
            self.lm_head_weight = embedding.wte.weight
        else:
            self.lm_head_weight = lm_head_weight
            if lm_head_weight is not None:
                lm_head_weight._qn_tied_duplicate = True
        self.n_embd = blocks[0].hidden_dim if len(blocks) else (config.n_embd if config else None)

    # ------------------------------------------------------------------ shapes
    def get_tensor_shapes(self, batch_size: int):
        T = getattr(self, "seq_len", None) or (self.config.n_positions if self.config else 1024)
        b = batch_size // group_size(self.tp_group) if self.sequence_parallel else batch_size
        return (b, T, self.n_embd)

    def _local_rows(self, t: torch.Tensor) -> torch.Tensor:
        """This TP rank's share of the micro-batch (sequence-parallel mode)."""
        n, r = group_size(self.tp_group), group_rank(self.tp_group)
        if t.shape[0] % n:
            raise ValueError(f"micro-batch {t.shape[0]} not divisible by tp={n} (sequence_parallel)")
        k = t.shape[0] // n
        return t[r * k:(r + 1) * k]

    # ----------------------------------------------------------------- forward
    def forward(self, x: torch.Tensor, position_ids: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None):
        if self.is_first_stage:
            if self.sequence_parallel:
                x = self._local_rows(x)
                if position_ids is not None:
                    position_ids = self._local_rows(position_ids)
            x = self.embedding(x, position_ids)
        pending = None                       # branch output whose dropout + residual add is still owed
        recompute = bool(getattr(self.config, "activation_checkpointing", False)) and self.training \
            and torch.is_grad_enabled()
        for blk in self.blocks:
            if recompute:
                pending, x = ops.checkpoint(blk.forward_fused, pending, x)
            else:
                pending, x = blk.forward_fused(pending, x)
        p_res = self.blocks[-1].mlp.resid_pdrop if len(self.blocks) else 0.0
        if not self.is_last_stage:
            return x if pending is None else ops.dropout_add(pending, x, p_res, self.training)
        if pending is None:
            x = ops.layer_norm(x, self.ln_f.weight, self.ln_f.bias, self.ln_f.eps)
        else:
            x, _ = ops.dropout_add_layer_norm(pending, x, self.ln_f.weight, self.ln_f.bias, self.ln_f.eps,
                                              p_res, self.training)
        if labels is None:
            if self.sequence_parallel:
                x = All_Gather.apply(x, self.tp_group, "reduce_scatter", 0)
            return ops.linear(x, self.lm_head_weight.to(x.dtype))
        return self.loss(x, labels)

    def loss(self, h: torch.Tensor, labels: torch.Tensor):
        """(sum of token losses, #supervised tokens) — fused lm_head + CE."""
        if self.shift_labels:
            shifted = torch.full_like(labels, -100)
            shifted[:, :-1] = labels[:, 1:]
            labels = shifted
        if self.sequence_parallel:
            labels = self._local_rows(labels)
        B, T, E = h.shape
        loss_sum, n = ops.linear_cross_entropy(h.reshape(B * T, E), self.lm_head_weight.to(h.dtype),
                                               labels.reshape(-1), -100)
        if self.sequence_parallel:
            stats = torch.stack([loss_sum.detach().float(), n.float()])
            dist.all_reduce(stats, group=self.tp_group)
            # value = global sum, gradient = local part only (summed over TP by the grad sync)
            loss_sum = loss_sum + (stats[0] - loss_sum.detach())
            n = stats[1]
        return loss_sum, n

    # -------------------------------------------------------------- grad syncs
    def sync_tied_weights_grad(self) -> None:
        """SUM the wte / lm_head gradient between first and last stage."""
        if self.is_first_stage and self.is_last_stage:
            return
        g = self.tied_group
        if g is None or group_size(g) < 2:
            return
        if self.is_first_stage and self.embedding is not None:
            w = self.embedding.wte.weight
        elif self.is_last_stage and self.lm_head_weight is not None:
            w = self.lm_head_weight
        else:
            return
        gw = grad_of(w)                      # fp32 main-grad bucket view when the buckets are high-precision
        if gw is None:
            w.grad = torch.zeros_like(w)
            gw = w.grad
        dist.all_reduce(gw, op=dist.ReduceOp.SUM, group=g)

    def sequence_parallel_params(self) -> List[nn.Parameter]:
        return [p for p in self.parameters()
                if not getattr(p, "_qn_tp_sharded", False) and not getattr(p, "_qn_tp_unique", False)]

    def sync_sequence_parallel_grads(self) -> None:
        if not self.sequence_parallel:
            return
        grads = [grad_of(p) for p in self.sequence_parallel_params() if grad_of(p) is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, group=self.tp_group)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    # ---------------------------------------------------------------- builders
    @classmethod
    def from_sharded_state_dict(cls, state_dict: dict, config, pp_rank: int, pp_size: int,
                                tp_rank: int, tp_size: int, tp_group, pp_group, device,
                                dtype=None, tied_group=None) -> "GPT2Stage":
        """Build from the per-rank dict of :func:`load_gpt2_distributed` (keys
        ``wte.weight``, ``wpe.weight``, ``h.{i}.…`` with GLOBAL layer indices,
        ``ln_f.*``)."""
        from ...parallel.pipeline_parallel.utils import distribute_layers
        is_first, is_last = pp_rank == 0, pp_rank == pp_size - 1
        layer_ids = distribute_layers(config.n_layer, pp_size, getattr(config, 'pp_layer_counts', None))[pp_rank]
        sd = state_dict
        embedding = None
        if is_first:
            embedding = GPT2Embedding(config.vocab_size, config.n_positions, config.n_embd,
                                      wte_weights=sd["wte.weight"], wpe_weights=sd.get("wpe.weight"),
                                      dropout_prob=config.embd_pdrop, device=device, dtype=dtype,
                                      use_wpe=not getattr(config, "use_rope", False))
        blocks = []
        for i in layer_ids:
            p = f"h.{i}"
            blocks.append(GPT2Block(
                config, tp_rank, tp_size, tp_group, device,
                ln_1_weight=sd[f"{p}.ln_1.weight"], ln_1_bias=sd[f"{p}.ln_1.bias"],
                ln_2_weight=sd[f"{p}.ln_2.weight"], ln_2_bias=sd[f"{p}.ln_2.bias"],
                attn_c_attn_weight=sd[f"{p}.attn.c_attn.weight"], attn_c_attn_bias=sd[f"{p}.attn.c_attn.bias"],
                attn_c_proj_weight=sd[f"{p}.attn.c_proj.weight"], attn_c_proj_bias=sd.get(f"{p}.attn.c_proj.bias"),
                mlp_c_fc_weight=sd[f"{p}.mlp.c_fc.weight"], mlp_c_fc_bias=sd[f"{p}.mlp.c_fc.bias"],
                mlp_c_proj_weight=sd[f"{p}.mlp.c_proj.weight"], mlp_c_proj_bias=sd.get(f"{p}.mlp.c_proj.bias"),
                dtype=dtype))
        ln_f, lm_head = None, None
        if is_last:
            ln_f = nn.LayerNorm(config.n_embd, eps=config.layer_norm_epsilon, device=device, dtype=dtype)
            with torch.no_grad():
                ln_f.weight.copy_(sd["ln_f.weight"])
                ln_f.bias.copy_(sd["ln_f.bias"])
            if not is_first:
                if "wte.weight" not in sd:
                    raise ValueError("wte.weight missing on the last stage (needed for the tied lm_head)")
                lm_head = nn.Parameter(sd["wte.weight"].detach().clone().to(device=device,
                                                                          dtype=dtype or sd["wte.weight"].dtype))
        return cls(embedding, nn.ModuleList(blocks), ln_f, lm_head, is_first, is_last, pp_group,
                   config=config, tp_group=tp_group, tied_group=tied_group)
