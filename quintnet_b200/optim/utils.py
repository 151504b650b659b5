"""Optimizer-side helpers: global gradient-norm clipping over the (tp, pp)
mesh axes (the reference clips with the local norm only — SURVEY §8-D15),
bucket-aware ``zero_grads`` and a linear-warmup LR schedule (the reference reads
``warmup_steps`` from YAML and ignores it — D16)."""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist

from ..parallel.data_parallel.components.bucket import grad_of


def _axis(pg_manager, name):
    if pg_manager is None:
        return None, 0, 1
    return pg_manager.get_group(name), pg_manager.axis_rank(name), pg_manager.axis_size(name)


def counts_for_norm(p: torch.Tensor, tp_rank: int, pp_size: int) -> bool:
    """A parameter replicated across TP is counted on tp-rank 0 only; the
    duplicate of a tied weight is not counted at all."""
    if getattr(p, "_qn_tied_duplicate", False) and pp_size > 1:
        return False
    if tp_rank != 0 and not getattr(p, "_qn_tp_sharded", False):
        return False
    return True


def global_grad_norm(parameters: Iterable[torch.Tensor], pg_manager=None) -> Optional[torch.Tensor]:
    tp_group, tp_rank, tp_size = _axis(pg_manager, "tp")
    pp_group, _, pp_size = _axis(pg_manager, "pp")
    grads = [grad_of(p) for p in parameters
             if grad_of(p) is not None and counts_for_norm(p, tp_rank, pp_size)]
    all_params_have_device = None
    if grads:
        all_params_have_device = grads[0].device
        norms = torch._foreach_norm(grads, 2.0)
        sq = torch.stack([n.float() for n in norms]).pow(2).sum()
    else:
        return None if pg_manager is None or (tp_size == 1 and pp_size == 1) else _zero_norm(pg_manager)
    if tp_size > 1:
        dist.all_reduce(sq, group=tp_group)
    if pp_size > 1:
        dist.all_reduce(sq, group=pp_group)
    return sq.sqrt()


def _zero_norm(pg_manager):
    dev = torch.device("cuda", torch.cuda.current_device()) if pg_manager.device_type == "cuda" \
        else torch.device("cpu")
    sq = torch.zeros((), device=dev)
    for name in ("tp", "pp"):
        g = pg_manager.get_group(name)
        if g is not None:
            dist.all_reduce(sq, group=g)
    return sq.sqrt()


def clip_grad_norm_global(parameters, max_norm: float, pg_manager=None) -> Optional[torch.Tensor]:
    params = [p for p in parameters if grad_of(p) is not None]
    norm = global_grad_norm(params, pg_manager)
    if norm is None or not params:
        return norm
    coef = (max_norm / (norm + 1e-6)).clamp(max=1.0)
    by_dtype = {}
    for p in params:
        g = grad_of(p)
        by_dtype.setdefault(g.dtype, []).append(g)
    for dt, gs in by_dtype.items():
        torch._foreach_mul_(gs, coef.to(dt))
    return norm


def zero_grads(model, optimizer=None) -> None:
    """Zero gradients without destroying DP bucket views."""
    if hasattr(model, "bucket_manager"):
        model.zero_grad()
        return
    if optimizer is not None and hasattr(optimizer, "zero_grad"):
        if getattr(optimizer, "owns_flat_grads", False):
            optimizer.zero_grad()
        else:
            optimizer.zero_grad(set_to_none=True)
    else:
        model.zero_grad(set_to_none=True)


class WarmupLR:
    """Linear warm-up to the base LR, then constant (or cosine to ``min_ratio``)."""

    def __init__(self, optimizer, warmup_steps: int = 0, total_steps: Optional[int] = None,
                 min_ratio: float = 1.0):
        self.optimizer, self.warmup_steps = optimizer, int(warmup_steps)
        self.total_steps, self.min_ratio = total_steps, min_ratio
        self.base_lrs = [g["lr"] for g in optimizer.param_groups]
        self.step_num = 0
        self._apply()

    def factor(self) -> float:
        import math
        s = self.step_num
        if self.warmup_steps > 0 and s < self.warmup_steps:
            return (s + 1) / self.warmup_steps
        if self.total_steps and self.min_ratio < 1.0:
            t = min(1.0, (s - self.warmup_steps) / max(1, self.total_steps - self.warmup_steps))
            return self.min_ratio + 0.5 * (1 - self.min_ratio) * (1 + math.cos(math.pi * t))
        return 1.0

    def _apply(self):
        f = self.factor()
        for g, b in zip(self.optimizer.param_groups, self.base_lrs):
            g["lr"] = b * f

    def step(self):
        self.step_num += 1
        self._apply()

    def state_dict(self):
        return {"step_num": self.step_num, "base_lrs": self.base_lrs}

    def load_state_dict(self, sd):
        self.step_num, self.base_lrs = sd["step_num"], sd["base_lrs"]
        self._apply()
