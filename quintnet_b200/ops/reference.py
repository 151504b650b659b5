"""Plain-PyTorch reference implementations of every op in :mod:`quintnet_b200.ops`.

These are (a) the CPU path used by the gloo plumbing tests and (b) the fp32
oracle the sm_100a kernels are tested against (tests/test_kernels_gpu.py).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


class _LinearAccumRef(torch.autograd.Function):
    """CPU twin of the fused weight-gradient accumulation: the backward ADDS dW into ``w.grad`` itself (when the
    buffer exists), reports it to the data-parallel bucket bookkeeping and returns no gradient for the weight."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.w_param, ctx.has_bias = w, b is not None
        return F.linear(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        from . import wgrad_done
        x, w = ctx.saved_tensors
        dy2, x2 = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
        dx = (dy2 @ w).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            dw = dy2.t() @ x2
            wp = ctx.w_param
            mg = getattr(wp, "main_grad", None)
            if mg is not None:                      # fp32 main-grad bucket view (low-precision parameters)
                mg.add_(dw.to(mg.dtype))
                wgrad_done(wp)
                dw = None
            elif wp.grad is not None:
                wp.grad.add_(dw.to(wp.grad.dtype))
                wgrad_done(wp)
                dw = None
        db = dy2.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db


def linear(x, w, b=None, act: Optional[str] = None):
    from . import fused_wgrad_accumulation
    if fused_wgrad_accumulation() and isinstance(w, torch.nn.Parameter) and torch.is_grad_enabled():
        y = _LinearAccumRef.apply(x, w, b)
    else:
        y = F.linear(x, w, b)
    if act == "gelu":
        y = F.gelu(y)
    elif act == "relu":
        y = F.relu(y)
    elif act is not None:
        raise ValueError(f"unknown activation {act}")
    return y


def fake_quant_rowwise_e4m3(t: torch.Tensor) -> torch.Tensor:
    """Row-wise amax-scaled e4m3 quantise -> dequantise (the numerics of the FP8 GEMM operands), with a
    straight-through gradient so backward stays in the wide format like the native op."""
    t2 = t.detach().float()
    amax = t2.abs().amax(dim=-1, keepdim=True)
    scale = torch.where(amax > 0, amax / 448.0, torch.ones_like(amax))
    q = (t2 / scale).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float() * scale
    return t + (q.to(t.dtype) - t).detach()


def mx_scales_e8m0(t: torch.Tensor, block: int = 32):
    """OCP MX shared exponents of a [..., K] tensor: one power-of-two scale per ``block`` elements along K,
    e = floor(log2(amax)) - 8 (8 = emax of e4m3), 2^-127 for an all-zero block.  Returns (biased uint8 [.., K/block],
    fp32 scale values)."""
    t2 = t.detach().float()
    K = t2.shape[-1]
    blk = t2.reshape(*t2.shape[:-1], K // block, block)
    amax = blk.abs().amax(dim=-1)
    e = torch.floor(torch.log2(amax.clamp(min=1e-38))) - 8.0
    # exact floor(log2) through frexp (log2 of a float can round up just below a power of two)
    m, ex = torch.frexp(amax)
    e = torch.where(amax > 0, ex.float() - 1.0 - 8.0, torch.full_like(amax, -127.0)).clamp(-127.0, 127.0)
    return (e + 127.0).to(torch.uint8), torch.exp2(e)


def fake_quant_mxfp8(t: torch.Tensor, block: int = 32) -> torch.Tensor:
    """MXFP8 quantise -> dequantise along the last dim (e4m3 elements x UE8M0 block scales): the numerics of the
    block-scaled tensor-core GEMM (``csrc/gemm_mxfp8.cu``), straight-through gradient."""
    t2 = t.detach().float()
    K = t2.shape[-1]
    _, scale = mx_scales_e8m0(t2, block)
    blk = t2.reshape(*t2.shape[:-1], K // block, block)
    q = (blk / scale.unsqueeze(-1)).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float() * scale.unsqueeze(-1)
    q = q.reshape(t2.shape)
    return t + (q.to(t.dtype) - t).detach()


def linear_fp8(x, w, b=None, act: Optional[str] = None, recipe: str = "rowwise"):
    if recipe == "mxfp8" and x.shape[-1] % 32 == 0:
        return linear(fake_quant_mxfp8(x), fake_quant_mxfp8(w), b, act)
    return linear(fake_quant_rowwise_e4m3(x), fake_quant_rowwise_e4m3(w), b, act)


def layer_norm(x, w, b, eps: float = 1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def add_layer_norm(x, residual, w, b, eps: float = 1e-5):
    """returns (LN(x + residual), x + residual)"""
    s = x + residual
    return F.layer_norm(s, (s.shape[-1],), w, b, eps), s


def dropout_add_layer_norm(x, residual, w, b, eps: float, p: float, training: bool):
    """returns (LN(h), h) with h = dropout(x) + residual"""
    h = dropout_add(x, residual, p, training)
    return F.layer_norm(h, (h.shape[-1],), w, b, eps), h


def gelu(x):
    return F.gelu(x)


def attention(q, k, v, causal: bool, dropout_p: float = 0.0, scale: Optional[float] = None,
              training: bool = True):
    """q,k,v: [B, H, T, d].  Explicit softmax(QK^T*scale)V."""
    d = q.shape[-1]
    scale = (1.0 / math.sqrt(d)) if scale is None else scale
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    if causal:
        Tq, Tk = s.shape[-2], s.shape[-1]
        mask = torch.ones(Tq, Tk, dtype=torch.bool, device=s.device).tril(diagonal=Tk - Tq)
        s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    if dropout_p > 0.0 and training:
        p = F.dropout(p, dropout_p, True)
    return torch.matmul(p, v.float()).to(q.dtype)


def cross_entropy_sum(logits, labels, ignore_index: int = -100):
    """returns (sum of token losses fp32, number of valid tokens)"""
    loss = F.cross_entropy(logits.float(), labels, ignore_index=ignore_index, reduction="sum")
    n = (labels != ignore_index).sum()
    return loss, n


def linear_cross_entropy(h, w, labels, ignore_index: int = -100):
    """h: [N, E], w: [V, E], labels: [N] -> (loss_sum, n_valid)"""
    return cross_entropy_sum(F.linear(h, w), labels, ignore_index)


def dropout_add(x, residual, p: float, training: bool):
    if p > 0.0 and training:
        x = F.dropout(x, p, True)
    return x + residual if residual is not None else x


def rope(x, cos, sin):
    """x: [B, H, T, d] (rotate-half convention); cos/sin: [T, d/2]."""
    d2 = x.shape[-1] // 2
    x1, x2 = x[..., :d2].float(), x[..., d2:].float()
    c, s = cos[None, None].float(), sin[None, None].float()
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1).to(x.dtype)


def rope_tables(T: int, d: int, base: float = 10000.0, device=None):
    inv = 1.0 / (base ** (torch.arange(0, d, 2, device=device, dtype=torch.float32) / d))
    ang = torch.arange(T, device=device, dtype=torch.float32)[:, None] * inv[None]
    return ang.cos(), ang.sin()


def adamw_step(p, g, m, v, step: int, lr: float, beta1: float, beta2: float, eps: float,
               weight_decay: float, grad_scale: float = 1.0, master=None):
    """Single-tensor AdamW (decoupled decay) in fp32; writes p/m/v in place.
    ``master`` is the optional fp32 master copy of a low-precision ``p``."""
    w = master if master is not None else p
    gf = g.float() * grad_scale
    m.mul_(beta1).add_(gf, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gf, gf, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v / bc2).sqrt_().add_(eps)
    wf = w.float()
    wf.mul_(1 - lr * weight_decay).addcdiv_(m / bc1, denom, value=-lr)
    w.copy_(wf)
    if master is not None:
        p.copy_(wf)
