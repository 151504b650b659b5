"""``FusedAdamW`` — AdamW over flat parameter/gradient buffers.

(The reference uses ``optim.Adam`` for ViT, ``trainer.py:89``, and ``optim.AdamW(weight_decay=0.01)`` for GPT-2,
``GPT2_Trainer.py:100-104``; its ``optimizers/`` package is docstring-only stubs, ``optimizers/__init__.py:1-12``.)

On B200 the whole update of a bucket is ONE launch of the hand-written sm_100a
kernel (K8, ``csrc/adamw.cu``): fp32 master/m/v, bf16 (or fp32) parameters,
optional gradient scale read from a device scalar (the clip coefficient, so no
host sync).  On CPU it runs the fp32 reference math.  With ``amsgrad=False,
weight_decay`` decoupled — numerically the same update as ``torch.optim.AdamW``
(``adam_mode=True`` drops the decoupled decay to L2-free Adam for the ViT
trainer, matching ``optim.Adam``).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from ..parallel.data_parallel.components.bucket import GradientBucket
from ..ops import optim_ops


class FusedAdamW(torch.optim.Optimizer):
    owns_flat_grads = True
    accepts_grad_scale = True          # step(grad_scale=<device scalar>) multiplies the gradients inside the kernel

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.01, buckets: Optional[List[GradientBucket]] = None,
                 master_weights: Optional[bool] = None, capacity_mb: int = 256,
                 grad_dtype: Optional[torch.dtype] = None):
        """``grad_dtype=torch.float32`` with bf16 parameters = fp32 MAIN-GRAD buckets: micro-batch gradients are
        accumulated in fp32 (by post-accumulate hooks, or directly by the wgrad GEMM epilogue — see
        ``ops.set_fused_wgrad_accumulation``) instead of in bf16 ``.grad`` tensors."""
        params = [p for p in params if p.requires_grad]
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._step = 0
        self._hooks = []
        if buckets is None:
            buckets = self._make_buckets(params, capacity_mb, grad_dtype)
            for b in buckets:
                if b.grad_dtype != b.param_dtype:      # nobody else owns these buckets: absorb at accumulate time
                    for j, p in enumerate(b.parameters):
                        self._hooks.append(p.register_post_accumulate_grad_hook(
                            lambda param, b=b, j=j: b.absorb(param, j)))
        self.buckets = buckets
        self.state_flat = []
        for b in self.buckets:
            assert b.flat_param is not None, "FusedAdamW needs buckets with flatten_parameters=True"
            low_precision = b.flat_param.dtype != torch.float32
            use_master = low_precision if master_weights is None else (master_weights and low_precision)
            self.state_flat.append({
                "m": torch.zeros(b.numel, dtype=torch.float32, device=b.device),
                "v": torch.zeros(b.numel, dtype=torch.float32, device=b.device),
                "master": b.flat_param.float().clone() if use_master else None,
            })

    @staticmethod
    def _make_buckets(params, capacity_mb, grad_dtype=None) -> List[GradientBucket]:
        groups = {}
        for p in params:
            groups.setdefault((p.device, p.dtype), []).append(p)
        return [GradientBucket(i, ps, grad_dtype if (grad_dtype is not None and ps[0].dtype != torch.float32)
                               else ps[0].dtype, flatten_parameters=True)
                for i, ps in enumerate(groups.values())]

    @classmethod
    def from_data_parallel(cls, ddp, **kw):
        bs = list(ddp.buckets.values())
        if any(b.flat_param is None for b in bs):
            raise ValueError("DataParallel must be built with BucketConfig(flatten_parameters=True)")
        return cls(list(ddp.parameters()), buckets=bs, **kw)

    def zero_grad(self, set_to_none: bool = False) -> None:
        for b in self.buckets:
            b.zero()
            b.attach_grads()

    @torch.no_grad()
    def step(self, closure=None, grad_scale: Optional[torch.Tensor] = None):
        loss = closure() if closure is not None else None
        self._step += 1
        g = self.param_groups[0]
        for b, st in zip(self.buckets, self.state_flat):
            for j, p in enumerate(b.parameters):
                b.absorb(p, j)
            optim_ops.adamw_flat(b.flat_param, b.flat_grad, st["m"], st["v"], st["master"],
                                 step=self._step, lr=g["lr"], beta1=g["betas"][0], beta2=g["betas"][1],
                                 eps=g["eps"], weight_decay=g["weight_decay"], grad_scale=grad_scale)
        return loss

    # checkpointing: flat state per bucket
    def state_dict(self):
        return {"step": self._step, "param_groups": [{k: v for k, v in g.items() if k != "params"}
                                                      for g in self.param_groups],
                "flat": [{k: (None if t is None else t.detach().cpu()) for k, t in st.items()}
                         for st in self.state_flat]}

    def load_state_dict(self, sd):
        self._step = sd["step"]
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            g.update(sg)
        for st, s in zip(self.state_flat, sd["flat"]):
            for k in st:
                if st[k] is not None and s[k] is not None:
                    st[k].copy_(s[k])
