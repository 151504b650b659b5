"""Optimizer math on flat buffers: native sm_100a kernel on CUDA (K8), fp32
reference on CPU."""
from __future__ import annotations

from typing import Optional

import torch

from . import native


def adamw_flat(p, g, m, v, master, *, step: int, lr: float, beta1: float, beta2: float, eps: float,
               weight_decay: float, grad_scale: Optional[torch.Tensor] = None,
               sqnorm_out: Optional[torch.Tensor] = None) -> None:
    """In-place AdamW on flat tensors.  ``grad_scale`` is an optional fp32 device
    scalar multiplied into the gradient (clip coefficient / loss-scale)."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    if p.is_cuda and native.op_enabled("adamw"):
        lib = native.require()
        pd = {torch.float32: 0, torch.bfloat16: 1}[p.dtype]
        gd = {torch.float32: 0, torch.bfloat16: 1}[g.dtype]
        if grad_scale is not None and grad_scale.dtype != torch.float32:
            grad_scale = grad_scale.float()
        rc = lib.qn_adamw(native.ptr(p), pd, native.ptr(g), gd, native.ptr(m), native.ptr(v),
                          native.ptr(master), native.ptr(grad_scale), p.numel(), lr, beta1, beta2, eps,
                          weight_decay, bc1, bc2, native.stream())
        native.check(rc, "qn_adamw")
        native.count()
        return
    gf = g.float()
    if grad_scale is not None:
        gf = gf * grad_scale.float()
    m.mul_(beta1).add_(gf, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(gf, gf, value=1.0 - beta2)
    w = master if master is not None else p
    wf = w.float()
    denom = (v / bc2).sqrt_().add_(eps)
    wf = wf * (1.0 - lr * weight_decay) - lr * (m / bc1) / denom
    w.copy_(wf)
    if master is not None:
        p.copy_(wf)
