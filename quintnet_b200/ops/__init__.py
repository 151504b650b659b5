"""Op layer: every hot op of the framework goes through here.

CUDA tensors are served by the hand-written sm_100a kernels in
``quintnet_b200/csrc`` (loaded from the in-tree ``lib/libquintnet_b200.so`` by
:mod:`.native`); CPU tensors by the plain-PyTorch references in
:mod:`.reference`.  There is no silent fallback on a GPU: if the native library
is missing the op raises, unless ``QUINTNET_ALLOW_FALLBACK=1`` is set
(debugging only).
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import reference
from . import native


def _use_native(name: str, *tensors) -> bool:
    t = tensors[0]
    if not t.is_cuda:
        return False
    return native.op_enabled(name)


# ----------------------------------------------------------------- linear
def linear(x, w, b=None, act: Optional[str] = None, fp8: bool = False):
    """y = act(x @ w.T + b);  w is [out, in] (nn.Linear layout).  ``fp8=True``: forward GEMM on e4m3 operands
    with row-wise (1 x K block) scales — fp8 tensor cores on B200, an exact emulation of the same quantisation
    on the reference path."""
    if _use_native("gemm", x) and x.dtype == torch.bfloat16:
        from . import cuda_ops
        if fp8 and _FP8_RECIPE == "mxfp8" and cuda_ops.mxfp8_ok(x.shape[-1], w.shape[0]) and \
                (b is None or b.dtype == torch.bfloat16):
            return cuda_ops.linear_mxfp8(x, w, b, act)            # block-scaled: tcgen05.mma.kind::mxf8f6f4.block_scale
        if fp8 and cuda_ops.fp8_ok(x.shape[-1], w.shape[0]):
            return cuda_ops.linear_fp8(x, w, b, act)              # row-wise scales: kind::f8f6f4 + epilogue dequant
        return cuda_ops.linear(x, w, b, act)
    if fp8:
        return reference.linear_fp8(x, w, b, act, recipe=_FP8_RECIPE)
    return reference.linear(x, w, b, act)


_FP8_RECIPE = os.environ.get("QUINTNET_FP8_RECIPE", "mxfp8")


def set_fp8_recipe(recipe: str) -> None:
    """``"mxfp8"`` (default): OCP MX block scaling — e4m3 elements, one UE8M0 scale per 32 elements along K, applied
    inside the tensor core (``tcgen05.mma.kind::mxf8f6f4.block_scale``); ``"rowwise"``: one fp32 scale per row,
    multiplied back in the GEMM epilogue (round-1 kernel)."""
    global _FP8_RECIPE
    if recipe not in ("mxfp8", "rowwise"):
        raise ValueError(f"unknown fp8 recipe {recipe!r}")
    _FP8_RECIPE = recipe


def fp8_recipe() -> str:
    return _FP8_RECIPE


def mlp(x, w1, b1, w2, b2):
    """gelu(x·W1ᵀ + b1)·W2ᵀ + b2 — one fused autograd op on B200 (GELU / GELU' live in GEMM epilogues)."""
    if _use_native("gemm", x):
        from . import cuda_ops
        if cuda_ops.mlp_ok(x, w1, b1, w2, b2):
            return cuda_ops.mlp(x, w1, b1, w2, b2)
    return linear(linear(x, w1, b1, "gelu"), w2, b2)


def layer_norm(x, w, b, eps: float = 1e-5):
    if _use_native("layernorm", x) and x.dtype in (torch.bfloat16, torch.float32):
        from . import cuda_ops
        return cuda_ops.layer_norm(x, w, b, eps)
    return reference.layer_norm(x, w, b, eps)


def add_layer_norm(x, residual, w, b, eps: float = 1e-5):
    if _use_native("layernorm", x) and x.dtype in (torch.bfloat16, torch.float32):
        from . import cuda_ops
        return cuda_ops.add_layer_norm(x, residual, w, b, eps)
    return reference.add_layer_norm(x, residual, w, b, eps)


def dropout_add_layer_norm(x, residual, w, b, eps: float, p: float, training: bool):
    """(LN(h), h), h = dropout(x, p) + residual — fused into one kernel per direction on B200."""
    if _use_native("layernorm", x) and _use_native("dropout", x):
        from . import cuda_ops
        if cuda_ops.dropout_add_layer_norm_ok(x):
            return cuda_ops.dropout_add_layer_norm(x, residual, w, b, eps, p if training else 0.0)
    return reference.dropout_add_layer_norm(x, residual, w.to(x.dtype), b.to(x.dtype), eps, p, training)


def gelu(x):
    if _use_native("gelu", x) and x.dtype == torch.bfloat16:
        from . import cuda_ops
        return cuda_ops.gelu(x)
    return reference.gelu(x)


def attention(q, k, v, causal: bool, dropout_p: float = 0.0, scale=None, training: bool = True,
              allow_library: bool = False):
    """q,k,v: [B, H, T, d] -> [B, H, T, d].  On CUDA this is the native shared-memory kernel for short sequences;
    shapes it cannot take raise unless the caller opts into the library path (``allow_library=True``, or
    ``QUINTNET_ALLOW_FALLBACK=1``) — no silent fallback."""
    if q.is_cuda:
        p = dropout_p if training else 0.0
        if _use_native("attention", q):
            from . import cuda_ops
            if cuda_ops.small_attention_ok(q, k, v, p):
                # short sequences (ViT: T = 50, d = 16): whole head in one CTA's shared memory, fwd and bwd one launch each
                return cuda_ops.small_attention(q, k, v, causal, scale)
            if not allow_library and os.environ.get("QUINTNET_ALLOW_FALLBACK", "0") != "1":
                raise RuntimeError(
                    f"no native attention kernel for q{tuple(q.shape)} {q.dtype} dropout={p}: the packed tcgen05 flash "
                    "kernel needs head_dim 64 (ops.attention_qkvpacked) and the shared-memory kernel needs bf16, "
                    "dropout 0 and T*T*8 + 4*T*(d+1)*4 <= 200 KB.  Set QUINTNET_ALLOW_FALLBACK=1 to use torch SDPA.")
        import torch.nn.functional as F
        return F.scaled_dot_product_attention(q, k, v, is_causal=causal, scale=scale, dropout_p=p)
    return reference.attention(q, k, v, causal, dropout_p, scale, training)


def attention_qkvpacked_ok(qkv, head_dim: int) -> bool:
    """True if the packed native flash-attention kernel (head_dim 64, bf16, CUDA) can take ``qkv``."""
    return (qkv.is_cuda and qkv.dtype == torch.bfloat16 and head_dim == 64
            and _use_native("attention", qkv))


def attention_qkvpacked(qkv, n_heads: int, causal: bool, dropout_p: float = 0.0, scale=None,
                        training: bool = True):
    """qkv: [B, T, 3*H*64] (layout [3, H, 64] in the last dim) -> [B, T, H*64]; native sm_100a kernel."""
    from . import cuda_ops
    return cuda_ops.attention_qkvpacked(qkv, n_heads, causal, dropout_p if training else 0.0, scale)


def cross_entropy_sum(logits, labels, ignore_index: int = -100):
    if _use_native("cross_entropy", logits):
        from . import cuda_ops
        return cuda_ops.cross_entropy_sum(logits, labels, ignore_index)
    return reference.cross_entropy_sum(logits, labels, ignore_index)


def linear_cross_entropy(h, w, labels, ignore_index: int = -100):
    """Fused lm_head + CE: never materialises fp32 [N, V] logits on GPU."""
    if _use_native("cross_entropy", h) and h.dtype == torch.bfloat16:
        from . import cuda_ops
        return cuda_ops.linear_cross_entropy(h, w, labels, ignore_index)
    return reference.linear_cross_entropy(h, w, labels, ignore_index)


def dropout_add(x, residual, p: float, training: bool):
    if _use_native("dropout", x) and x.dtype == torch.bfloat16 and p > 0.0 and training:
        from . import cuda_ops
        return cuda_ops.dropout_add(x, residual, p)
    return reference.dropout_add(x, residual, p, training)


def rope(x, cos, sin):
    if _use_native("rope", x) and x.dtype == torch.bfloat16:
        from . import cuda_ops
        return cuda_ops.rope(x, cos, sin)
    return reference.rope(x, cos, sin)


rope_tables = reference.rope_tables


# ------------------------------------------------------------------ activation checkpointing
def dropout_rng_state():
    """(seed, offset) of the native dropout counter (the kernels regenerate their masks from it), or None."""
    from . import native
    if not native.available():
        return None
    from . import cuda_ops
    return cuda_ops._PhiloxState.seed, cuda_ops._PhiloxState.offset


def set_dropout_rng_state(state) -> None:
    if state is not None:
        from . import cuda_ops
        cuda_ops._PhiloxState.seed, cuda_ops._PhiloxState.offset = state


class _ReplayDropoutRng:
    """Recomputation must draw the SAME dropout masks as the original forward: rewind the native counter to where
    the checkpointed region started, and put it back afterwards (torch's own generators are handled by
    ``preserve_rng_state``)."""

    def __init__(self, state):
        self.state, self.saved = state, None

    def __enter__(self):
        self.saved = dropout_rng_state()
        set_dropout_rng_state(self.state)

    def __exit__(self, *exc):
        set_dropout_rng_state(self.saved)
        return False


def checkpoint(fn, *args):
    """Run ``fn(*args)`` without keeping its intermediate activations; they are recomputed in backward."""
    import contextlib
    from torch.utils.checkpoint import checkpoint as _ckpt
    state = dropout_rng_state()

    def context_fn():
        return contextlib.nullcontext(), _ReplayDropoutRng(state)

    return _ckpt(fn, *args, use_reentrant=False, context_fn=context_fn, preserve_rng_state=True)


# ------------------------------------------------------------------ fused weight-gradient accumulation
_FUSE_WGRAD_ACCUM = os.environ.get("QUINTNET_FUSE_WGRAD_ACCUM", "0") == "1"


def set_fused_wgrad_accumulation(on: bool) -> None:
    """When on, linear layers whose weight already has a ``.grad`` buffer ADD their weight gradient into it from
    inside the backward op (on B200: in the wgrad GEMM's epilogue) and hand autograd no gradient for the weight —
    one pass over the gradient instead of GEMM-store + elementwise accumulate.  Off by default."""
    global _FUSE_WGRAD_ACCUM
    _FUSE_WGRAD_ACCUM = bool(on)


def fused_wgrad_accumulation() -> bool:
    return _FUSE_WGRAD_ACCUM


def wgrad_done(w) -> None:
    """Nothing to do: autograd still runs the weight's AccumulateGrad node (with an undefined gradient) after an op
    that returned ``None`` for it, and fires the post-accumulate hooks the data-parallel buckets count on — after the
    op's backward, i.e. after the in-place accumulation.  (``tests/test_gpt2.py`` runs the 3-D suites with the switch on.)"""
    return None
