"""autograd wrappers around the hand-written sm_100a kernels (ctypes → C ABI).

Every function here has a plain-PyTorch twin in :mod:`.reference`; the GPU
tests compare the two (fp32 oracle).  What each op replaces in the reference (SURVEY §2.9 ids): ``linear`` / ``mlp``
= the ``nn.Linear`` inside Column/RowParallelLinear (G1-G4: ``parallelism/tensor_parallel/layers.py:119,211``,
``utils/GPT2/gpt2_mlp.py:143-155``); ``attention_qkvpacked`` = ``F.scaled_dot_product_attention`` (G6,
``utils/GPT2/gpt2_attention.py:156-161``); ``layer_norm`` family = ``nn.LayerNorm`` (G7, ``gpt2_block.py:175,184``);
``linear_cross_entropy`` = lm_head matmul + ``CrossEntropyLoss`` (G5 + G9, ``gpt2_stage.py:108``, ``schedule.py:325-328``).  Layout conventions follow the kernels:
row-major contiguous bf16, leading dimensions multiples of 8 elements.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn.functional as F

from . import native

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_MUL_AUX = 0, 1, 2, 3
_DT = {torch.float32: 0, torch.bfloat16: 1}


def _lib():
    return native.require()


# --------------------------------------------------------------------- GEMM
def gemm_ok(M: int, N: int, K: int, *lds) -> bool:
    return M > 0 and N > 0 and K > 0 and all(int(l) % 8 == 0 for l in lds)


def gemm(a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, *, a_mn: bool = False,
         b_mn: bool = False, out: Optional[torch.Tensor] = None, ldc: Optional[int] = None,
         epilogue: int = EPI_NONE, bias: Optional[torch.Tensor] = None,
         z: Optional[torch.Tensor] = None, block_n: int = 0, max_ctas: int = 0,
         split_k: int = 0) -> torch.Tensor:
    """C[M,N] = epilogue(A·B) on tcgen05.  ``a``/``b`` are 2-D row-major bf16 tensors
    (possibly with a row stride); ``a_mn`` / ``b_mn`` say the tensor is stored
    ``[K, M]`` / ``[K, N]`` instead of ``[M, K]`` / ``[N, K]``."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1
    lda, ldb = a.stride(0), b.stride(0)
    if out is None:
        ldc = ldc or N
        out = torch.empty((M, ldc), dtype=torch.bfloat16, device=a.device)
        out_view = out if ldc == N else out[:, :N]
    else:
        ldc = out.stride(0)
        out_view = out
    ldz = z.stride(0) if z is not None else 0
    splits = _auto_splits(M, N, K) if (epilogue == EPI_NONE and split_k != 1) else 1
    if split_k > 1 and epilogue == EPI_NONE:
        splits = split_k
    if splits > 1 and N % 4 == 0:
        acc32 = torch.zeros((M, N), dtype=torch.float32, device=a.device)
        rc = _lib().qn_gemm_bf16_splitk(a.data_ptr(), b.data_ptr(), out.data_ptr(), acc32.data_ptr(), M, N, K,
                                        lda, ldb, ldc, int(a_mn), int(b_mn), splits, block_n, native.stream())
        native.check(rc, f"qn_gemm_bf16_splitk(M={M},N={N},K={K},splits={splits})")
        native.count(3)
        return out_view
    rc = _lib().qn_gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, lda, ldb, ldc,
                             int(a_mn), int(b_mn), epilogue, native.ptr(bias), native.ptr(z), ldz,
                             block_n, max_ctas, native.stream())
    native.check(rc, f"qn_gemm_bf16(M={M},N={N},K={K},a_mn={a_mn},b_mn={b_mn})")
    native.count()
    return out_view


def _auto_splits(M: int, N: int, K: int) -> int:
    """Split-K factor for few-tile / long-K problems (weight gradients): fill ~148 SMs."""
    bn = 256 if N > 128 else 128
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    num_kb = (K + 63) // 64
    if tiles >= 100 or num_kb < 32:
        return 1
    s = min(148 // tiles, num_kb // 8, 16)
    return s if s >= 2 else 1


def wgrad(w_param, dy2: torch.Tensor, x2: torch.Tensor, N: int, K: int, M: int):
    """dW[N, K] = dYᵀ[N, M] · X[M, K] (both operands stored [M, ·] -> MN-major).  Returns the gradient for autograd —
    or ``None`` after ADDING it into ``w_param.grad`` inside the GEMM epilogue when fused weight-gradient
    accumulation is on and the gradient buffer exists (micro-batches after the first)."""
    from . import fused_wgrad_accumulation, wgrad_done
    mg = getattr(w_param, "main_grad", None) if w_param is not None else None
    if (fused_wgrad_accumulation() and mg is not None and mg.dtype == torch.float32 and mg.dim() == 2
            and tuple(mg.shape) == (N, K) and mg.stride(1) == 1 and mg.stride(0) % 4 == 0 and K % 4 == 0
            and mg.data_ptr() % 16 == 0):
        # fp32 main-grad bucket view: the GEMM epilogue red.adds its accumulator tiles straight into it
        rc = _lib().qn_gemm_bf16_wgrad_accum_f32(dy2.data_ptr(), x2.data_ptr(), mg.data_ptr(), N, K, M, dy2.stride(0),
                                                 x2.stride(0), mg.stride(0), _auto_splits(N, K, M), 0, native.stream())
        native.check(rc, f"qn_gemm_bf16_wgrad_accum_f32(N={N},K={K},M={M})")
        native.count()
        wgrad_done(w_param)
        return None
    g = getattr(w_param, "grad", None) if w_param is not None else None
    if (fused_wgrad_accumulation() and g is not None and g.dtype == torch.bfloat16 and g.dim() == 2
            and tuple(g.shape) == (N, K) and g.stride(1) == 1 and g.stride(0) % 8 == 0 and K % 4 == 0
            and g.data_ptr() % 16 == 0):
        splits = _auto_splits(N, K, M)
        acc32 = torch.zeros((N, K), dtype=torch.float32, device=g.device) if splits > 1 else None
        rc = _lib().qn_gemm_bf16_wgrad_accum(dy2.data_ptr(), x2.data_ptr(), g.data_ptr(), native.ptr(acc32), N, K, M,
                                             dy2.stride(0), x2.stride(0), g.stride(0), splits, 0, native.stream())
        native.check(rc, f"qn_gemm_bf16_wgrad_accum(N={N},K={K},M={M})")
        native.count(3 if splits > 1 else 1)
        wgrad_done(w_param)
        return None
    return gemm(dy2, x2, N, K, M, a_mn=True, b_mn=True)


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act):
        K = x.shape[-1]
        N = w.shape[0]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        wq = w if w.dtype == torch.bfloat16 else w.to(torch.bfloat16)
        z = None
        if act == "gelu":
            z = torch.empty((M, N), dtype=torch.bfloat16, device=x.device) if ctx.needs_input_grad[0] or \
                ctx.needs_input_grad[1] else None
            y = gemm(x2, wq, M, N, K, epilogue=EPI_BIAS_GELU, bias=b, z=z)
        elif act is None:
            y = gemm(x2, wq, M, N, K, epilogue=EPI_BIAS if b is not None else EPI_NONE, bias=b)
        else:
            raise ValueError(f"native linear: unsupported activation {act}")
        ctx.save_for_backward(x2, wq, z)
        ctx.act, ctx.has_bias, ctx.xshape = act, b is not None, x.shape
        ctx.wdtype = w.dtype
        ctx.w_param = w if w.dtype == torch.bfloat16 else None      # target of fused gradient accumulation
        ctx.b_param = b
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w, z = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        dy2 = dy.reshape(M, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        if ctx.act == "gelu":
            dy2 = dy2 * z            # z holds gelu'(pre-activation), written by the forward epilogue
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dX[M,K] = dY[M,N] · W[N,K]        (B stored [K_red=N, N_out=K] -> MN-major)
            dx = gemm(dy2, w, M, K, N, b_mn=True).view(ctx.xshape)
        if ctx.needs_input_grad[1]:
            # dW[N,K] = dYᵀ[N,M] · X[M,K]       (both operands stored [K_red=M, ·] -> MN-major)
            dw = wgrad(getattr(ctx, "w_param", None), dy2, x2, N, K, M)
            if dw is not None and ctx.wdtype != torch.bfloat16:
                dw = dw.to(ctx.wdtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = bias_grad(getattr(ctx, "b_param", None), dy2)
        return dx, dw, db, None


def quantize_rowwise(x2: torch.Tensor):
    """bf16 [R, C] -> (e4m3 bytes [R, C] as uint8, fp32 dequantisation scale [R]); one scale per row
    (a 1 x C block), amax-based, saturating."""
    R, C = x2.shape
    q = torch.empty((R, C), dtype=torch.uint8, device=x2.device)
    sc = torch.empty((R,), dtype=torch.float32, device=x2.device)
    rc = native.require().qn_quantize_rowwise_e4m3(x2.data_ptr(), q.data_ptr(), sc.data_ptr(), R, C, x2.stride(0), C,
                                                   native.stream())
    native.check(rc, "qn_quantize_rowwise_e4m3")
    native.count()
    return q, sc


def fp8_ok(K: int, N: int) -> bool:
    return K % 16 == 0 and N % 8 == 0


def gemm_fp8(a8: torch.Tensor, sa: torch.Tensor, b8: torch.Tensor, sb: torch.Tensor, M: int, N: int, K: int, *,
             epilogue: int = EPI_NONE, bias: Optional[torch.Tensor] = None, z: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C[M, N] (bf16) = epilogue((A8 · B8ᵀ) * sa[:, None] * sb[None, :]) on the fp8 tensor cores
    (tcgen05.mma kind::f8f6f4, fp32 accumulation in TMEM)."""
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a8.device)
    rc = native.require().qn_gemm_fp8(a8.data_ptr(), b8.data_ptr(), out.data_ptr(), M, N, K, a8.stride(0), b8.stride(0),
                                      out.stride(0), sa.data_ptr(), sb.data_ptr(), epilogue, native.ptr(bias),
                                      native.ptr(z), z.stride(0) if z is not None else 0, 0, native.stream())
    native.check(rc, f"qn_gemm_fp8(M={M},N={N},K={K})")
    native.count()
    return out


class _LinearFP8(torch.autograd.Function):
    """Forward in FP8 (e4m3 activations and weights, row-wise scales, fp32 accumulation); backward in bf16 from
    the saved bf16 operands — the usual mixed recipe (the gradient GEMMs keep the wider format)."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        K, N = x.shape[-1], w.shape[0]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        wq = w if w.dtype == torch.bfloat16 else w.to(torch.bfloat16)
        a8, sa = quantize_rowwise(x2)
        b8, sb = quantize_rowwise(wq)
        z = None
        if act == "gelu":
            z = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
            y = gemm_fp8(a8, sa, b8, sb, M, N, K, epilogue=EPI_BIAS_GELU, bias=b, z=z)
        elif act is None:
            y = gemm_fp8(a8, sa, b8, sb, M, N, K, epilogue=EPI_BIAS if b is not None else EPI_NONE, bias=b)
        else:
            raise ValueError(f"native fp8 linear: unsupported activation {act}")
        ctx.save_for_backward(x2, wq, z)
        ctx.act, ctx.has_bias, ctx.xshape, ctx.wdtype = act, b is not None, x.shape, w.dtype
        ctx.w_param = w if w.dtype == torch.bfloat16 else None
        ctx.b_param = b
        return y.view(*x.shape[:-1], N)

    backward = _Linear.backward


def linear_fp8(x, w, b=None, act: Optional[str] = None):
    return _LinearFP8.apply(x, w, b, act)


# ------------------------------------------------------------------ block-scaled FP8 (MXFP8, tcgen05 .block_scale)
def mxfp8_ok(K: int, N: int) -> bool:
    return K % 128 == 0 and N % 8 == 0


def quantize_mx(x2: torch.Tensor):
    """bf16 [R, K] -> (e4m3 bytes [R, K], UE8M0 scale atoms [ceil(R/128) * K/128 * 512]) — ``csrc/gemm_mxfp8.cu``."""
    R, K = x2.shape
    assert x2.dtype == torch.bfloat16 and x2.stride(1) == 1 and K % 128 == 0
    q = torch.empty((R, K), dtype=torch.uint8, device=x2.device)
    nblk = (R + 127) // 128
    sf = (torch.zeros if R % 128 else torch.empty)(nblk * (K // 128) * 512, dtype=torch.uint8, device=x2.device)
    rc = _lib().qn_quantize_mxfp8(x2.data_ptr(), q.data_ptr(), sf.data_ptr(), R, K, x2.stride(0), K, native.stream())
    native.check(rc, f"qn_quantize_mxfp8(R={R},K={K})")
    native.count()
    return q, sf


def gemm_mxfp8(a8, sfa, b8, sfb, M: int, N: int, K: int, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C[M, N] (bf16) = (A8 ⊙ SFA)·(B8 ⊙ SFB)ᵀ (+ bias) on ``tcgen05.mma.kind::mxf8f6f4.block_scale``: the UE8M0 scales
    go shared memory -> tensor memory (tcgen05.cp) and are applied by the tensor core per 32-element K block."""
    out = torch.empty((M, N), dtype=torch.bfloat16, device=a8.device)
    rc = _lib().qn_gemm_mxfp8(a8.data_ptr(), sfa.data_ptr(), b8.data_ptr(), sfb.data_ptr(), out.data_ptr(), M, N, K,
                              a8.stride(0), b8.stride(0), N, native.ptr(bias), native.stream())
    native.check(rc, f"qn_gemm_mxfp8(M={M},N={N},K={K})")
    native.count()
    return out


class _LinearMXFP8(torch.autograd.Function):
    """Forward on the block-scaled FP8 tensor cores (MXFP8 activations AND weights, fp32 accumulation); backward in bf16
    from the saved bf16 operands (dgrad / wgrad keep the wide format — the usual mixed recipe).  GELU, when requested,
    runs as the stand-alone fused kernel on the GEMM output (its derivative is taken from the saved pre-activation)."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        K, N = x.shape[-1], w.shape[0]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        wq = w if w.dtype == torch.bfloat16 else w.to(torch.bfloat16)
        a8, sfa = quantize_mx(x2)
        b8, sfb = quantize_mx(wq if wq.is_contiguous() else wq.contiguous())
        y = gemm_mxfp8(a8, sfa, b8, sfb, M, N, K, b)
        z = None
        if act == "gelu":
            pre = y
            y = torch.nn.functional.gelu(pre.float()).to(torch.bfloat16)
            u = pre.float()
            z = (0.5 * (1.0 + torch.erf(u * 0.7071067811865476)) + u * torch.exp(-0.5 * u * u) * 0.3989422804014327).to(torch.bfloat16)
        elif act is not None:
            raise ValueError(f"native mxfp8 linear: unsupported activation {act}")
        ctx.save_for_backward(x2, wq, z)
        ctx.act, ctx.has_bias, ctx.xshape, ctx.wdtype = act, b is not None, x.shape, w.dtype
        ctx.w_param = w if w.dtype == torch.bfloat16 else None
        ctx.b_param = b
        return y.view(*x.shape[:-1], N)

    backward = _Linear.backward


def linear_mxfp8(x, w, b=None, act: Optional[str] = None):
    return _LinearMXFP8.apply(x, w, b, act)


class _FusedMLP(torch.autograd.Function):
    """y = (gelu(x·W1ᵀ + b1))·W2ᵀ + b2 with the GELU fused into the first GEMM's epilogue (forward) and
    gelu' fused into the dgrad GEMM's epilogue (backward): no stand-alone activation kernels."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        E, I = x.shape[-1], w1.shape[0]
        x2 = x.reshape(-1, E)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        z = torch.empty((M, I), dtype=torch.bfloat16, device=x.device)
        a = gemm(x2, w1, M, I, E, epilogue=EPI_BIAS_GELU, bias=b1, z=z)
        y = gemm(a, w2, M, w2.shape[0], I, epilogue=EPI_BIAS if b2 is not None else EPI_NONE, bias=b2)
        ctx.save_for_backward(x2, w1, w2, z, a)
        ctx.xshape, ctx.has_b1, ctx.has_b2 = x.shape, b1 is not None, b2 is not None
        ctx.b1_param, ctx.b2_param = b1, b2
        return y.view(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w1, w2, z, a = ctx.saved_tensors
        M, E = x2.shape
        I, O = w1.shape[0], w2.shape[0]
        dy2 = dy.reshape(M, O)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dz = gemm(dy2, w2, M, I, O, b_mn=True, epilogue=EPI_MUL_AUX, z=z)      # (dY·W2) ⊙ gelu'(z), z := gelu'
        dw2 = wgrad(w2, dy2, a, O, I, M) if ctx.needs_input_grad[3] else None
        db2 = bias_grad(ctx.b2_param, dy2) if (ctx.has_b2 and ctx.needs_input_grad[4]) else None
        dx = gemm(dz, w1, M, E, I, b_mn=True).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        dw1 = wgrad(w1, dz, x2, I, E, M) if ctx.needs_input_grad[1] else None
        db1 = bias_grad(ctx.b1_param, dz) if (ctx.has_b1 and ctx.needs_input_grad[2]) else None
        return dx, dw1, db1, dw2, db2


def mlp(x, w1, b1, w2, b2):
    return _FusedMLP.apply(x, w1, b1, w2, b2)


def mlp_ok(x, w1, b1, w2, b2) -> bool:
    E, I, O = x.shape[-1], w1.shape[0], w2.shape[0]
    M = x.numel() // E
    ts = [t for t in (x, w1, b1, w2, b2) if t is not None]
    return (all(t.dtype == torch.bfloat16 and t.is_cuda for t in ts) and gemm_ok(M, I, E, E, I, O)
            and all(t.data_ptr() % 16 == 0 for t in ts) and M % 8 == 0)


# Bias gradients that a producer kernel already computed (the fused LayerNorm backward sums the columns of the gradient
# it hands to the branch): keyed by the gradient tensor's address, valid only while that very tensor object is alive
# (weak reference) — a freed-and-reused address can never hit.
import weakref

_COLSUM_CACHE = {}


def offer_colsum(grad: torch.Tensor, cs: torch.Tensor) -> None:
    if len(_COLSUM_CACHE) > 64:
        _COLSUM_CACHE.clear()
    _COLSUM_CACHE[grad.data_ptr()] = (weakref.ref(grad), cs, grad.numel())


def colsum(x2: torch.Tensor) -> torch.Tensor:
    ent = _COLSUM_CACHE.pop(x2.data_ptr(), None)
    if ent is not None:
        ref, cs, n = ent
        if ref() is not None and n == x2.numel() and cs.numel() == x2.shape[-1]:
            return cs
    rows, cols = x2.shape
    acc = torch.empty(cols, dtype=torch.float32, device=x2.device)
    out = torch.empty(cols, dtype=torch.bfloat16, device=x2.device)
    native.check(_lib().qn_colsum(x2.data_ptr(), acc.data_ptr(), out.data_ptr(), rows, cols, x2.stride(0),
                                  native.stream()), "qn_colsum")
    native.count(3)
    return out


def bias_grad(b_param, dy2: torch.Tensor):
    """Bias gradient = column sums of dY.  With fp32 main grads and fused accumulation the sums are accumulated straight
    into ``b_param.main_grad`` by the kernel (one launch, nothing for autograd to add) and ``None`` is returned."""
    from . import fused_wgrad_accumulation, wgrad_done
    mg = getattr(b_param, "main_grad", None) if b_param is not None else None
    if (fused_wgrad_accumulation() and mg is not None and mg.dtype == torch.float32 and mg.dim() == 1
            and mg.numel() == dy2.shape[-1] and mg.is_contiguous() and dy2.data_ptr() not in _COLSUM_CACHE):
        rows, cols = dy2.shape
        native.check(_lib().qn_colsum_accum_f32(dy2.data_ptr(), mg.data_ptr(), rows, cols, dy2.stride(0), native.stream()),
                     "qn_colsum_accum_f32")
        native.count()
        wgrad_done(b_param)
        return None
    return colsum(dy2)


def linear(x, w, b=None, act: Optional[str] = None):
    K, N = x.shape[-1], w.shape[0]
    M = x.numel() // K
    bias_ok = b is None or (b.dtype == torch.bfloat16 and N % 8 == 0)
    if not gemm_ok(M, N, K, K, N) or not bias_ok or act not in (None, "gelu") or \
            (x.data_ptr() % 16) or (w.data_ptr() % 16):
        # shapes the TMA path cannot express (e.g. the 10-class ViT head): library GEMM
        from . import reference
        return reference.linear(x, w.to(x.dtype), None if b is None else b.to(x.dtype), act)
    return _Linear.apply(x, w, b, act)


# ---------------------------------------------------------------- LayerNorm
_LN_PARTS = 148 * 4


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps, residual):
        cols = x.shape[-1]
        x2 = x.reshape(-1, cols)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        rows = x2.shape[0]
        res2 = None
        if residual is not None:
            res2 = residual.reshape(-1, cols)
            if not res2.is_contiguous():
                res2 = res2.contiguous()
        y = torch.empty_like(x2)
        s = torch.empty_like(x2) if residual is not None else None
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        wq, bq = w.to(x.dtype), b.to(x.dtype)
        rc = _lib().qn_layernorm_fwd(x2.data_ptr(), native.ptr(res2), wq.data_ptr(), bq.data_ptr(),
                                     y.data_ptr(), native.ptr(s), mean.data_ptr(), rstd.data_ptr(), rows, cols,
                                     float(eps), _DT[x.dtype], native.stream())
        native.check(rc, "qn_layernorm_fwd")
        native.count()
        ctx.save_for_backward(s if s is not None else x2, wq, mean, rstd)
        ctx.shape, ctx.has_res, ctx.wdtype = x.shape, residual is not None, w.dtype
        if residual is not None:
            return y.view(x.shape), s.view(x.shape)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy, dsum=None):
        xin, w, mean, rstd = ctx.saved_tensors
        rows, cols = xin.shape
        dy2 = dy.reshape(rows, cols)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dadd = None
        if ctx.has_res and dsum is not None:
            dadd = dsum.reshape(rows, cols)
            if not dadd.is_contiguous():
                dadd = dadd.contiguous()
        dx = torch.empty_like(xin)
        dg = torch.empty(cols, dtype=xin.dtype, device=xin.device)
        db = torch.empty(cols, dtype=xin.dtype, device=xin.device)
        nparts = min(_LN_PARTS, (rows + 3) // 4)
        ws = torch.empty(2 * nparts * cols, dtype=torch.float32, device=xin.device)
        rc = _lib().qn_layernorm_bwd(dy2.data_ptr(), xin.data_ptr(), w.data_ptr(), mean.data_ptr(),
                                     rstd.data_ptr(), native.ptr(dadd), dx.data_ptr(), dg.data_ptr(),
                                     db.data_ptr(), ws.data_ptr(), nparts, rows, cols, _DT[xin.dtype],
                                     native.stream())
        native.check(rc, "qn_layernorm_bwd")
        native.count(2)
        dxv = dx.view(ctx.shape)
        return dxv, dg.to(ctx.wdtype), db.to(ctx.wdtype), None, (dxv if ctx.has_res else None)


def _ln_ok(x, cols):
    vn = 8 if x.dtype == torch.bfloat16 else 4
    return cols % vn == 0 and cols <= 256 * vn


def layer_norm(x, w, b, eps=1e-5):
    if not _ln_ok(x, x.shape[-1]):
        from . import reference
        return reference.layer_norm(x, w.to(x.dtype), b.to(x.dtype), eps)
    return _LayerNorm.apply(x, w, b, eps, None)


def add_layer_norm(x, residual, w, b, eps=1e-5):
    if not _ln_ok(x, x.shape[-1]):
        from . import reference
        return reference.add_layer_norm(x, residual, w.to(x.dtype), b.to(x.dtype), eps)
    return _LayerNorm.apply(x, w, b, eps, residual)


# --------------------------------------------------------------------- GELU
class _Gelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        xc = x.contiguous()
        y = torch.empty_like(xc)
        native.check(_lib().qn_gelu_fwd(xc.data_ptr(), y.data_ptr(), xc.numel(), native.stream()), "qn_gelu_fwd")
        native.count()
        ctx.save_for_backward(xc)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dyc = dy.contiguous()
        dx = torch.empty_like(x)
        native.check(_lib().qn_gelu_bwd(dyc.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel(), native.stream()),
                     "qn_gelu_bwd")
        native.count()
        return dx


def gelu(x):
    if x.numel() % 8:
        return F.gelu(x)
    return _Gelu.apply(x)


# ------------------------------------------------------------ dropout + add
class _PhiloxState:
    seed = None
    offset = 0

    @classmethod
    def next(cls, n_elems: int):
        if cls.seed is None:
            cls.seed = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
        off = cls.offset
        cls.offset += 2 * ((n_elems + 7) // 8)
        return cls.seed, off


def manual_seed(seed: int) -> None:
    _PhiloxState.seed, _PhiloxState.offset = int(seed) & 0x7FFFFFFFFFFFFFFF, 0


class _DropoutAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, p):
        xc = x.contiguous()
        rc_ = residual.contiguous() if residual is not None else None
        y = torch.empty_like(xc)
        seed, off = _PhiloxState.next(xc.numel())
        native.check(_lib().qn_dropout_add_fwd(xc.data_ptr(), native.ptr(rc_), y.data_ptr(), xc.numel(), float(p),
                                               seed, off, native.stream()), "qn_dropout_add_fwd")
        native.count()
        ctx.p, ctx.seed, ctx.off, ctx.has_res = float(p), seed, off, residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        dyc = dy.contiguous()
        dx = torch.empty_like(dyc)
        native.check(_lib().qn_dropout_bwd(dyc.data_ptr(), dx.data_ptr(), dyc.numel(), ctx.p, ctx.seed, ctx.off,
                                           native.stream()), "qn_dropout_bwd")
        native.count()
        return dx, (dy if ctx.has_res else None), None


def dropout_add(x, residual, p: float):
    if x.numel() % 8:
        from . import reference
        return reference.dropout_add(x, residual, p, True)
    return _DropoutAdd.apply(x, residual, p)


class _DropoutAddLayerNorm(torch.autograd.Function):
    """(LN(h), h) with h = dropout(x)·1/(1-p) + residual — ONE kernel forward, ONE backward
    (LN', residual-gradient add and dropout' fused; the mask is regenerated from the Philox counter)."""

    @staticmethod
    def forward(ctx, x, residual, w, b, eps, p):
        cols = x.shape[-1]
        x2 = x.reshape(-1, cols)
        r2 = residual.reshape(-1, cols)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        if not r2.is_contiguous():
            r2 = r2.contiguous()
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        h = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        wq, bq = w.to(x.dtype), b.to(x.dtype)
        seed, off = _PhiloxState.next(x2.numel()) if p > 0.0 else (0, 0)
        rc = _lib().qn_dropout_add_layernorm_fwd(x2.data_ptr(), r2.data_ptr(), wq.data_ptr(), bq.data_ptr(),
                                                 y.data_ptr(), h.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                 rows, cols, float(eps), float(p), seed, off, native.stream())
        native.check(rc, "qn_dropout_add_layernorm_fwd")
        native.count()
        ctx.save_for_backward(h, wq, mean, rstd)
        ctx.args = (float(p), seed, off, x.shape, w.dtype)
        ctx.ln_params = (w, b)                      # targets of in-place fp32 main-grad accumulation (if they have one)
        return y.view(x.shape), h.view(x.shape)

    @staticmethod
    def backward(ctx, dy, dh):
        h, w, mean, rstd = ctx.saved_tensors
        p, seed, off, shape, wdtype = ctx.args
        rows, cols = h.shape
        dy2 = dy.reshape(rows, cols)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dh2 = None
        if dh is not None:
            dh2 = dh.reshape(rows, cols)
            if not dh2.is_contiguous():
                dh2 = dh2.contiguous()
        dres = torch.empty_like(h)
        dx = torch.empty_like(h) if p > 0.0 else None
        dg = torch.empty(cols, dtype=h.dtype, device=h.device)
        db = torch.empty(cols, dtype=h.dtype, device=h.device)
        nparts = min(_LN_PARTS, (rows + 3) // 4)
        ws = torch.empty(max(2 * nparts, 3) * cols, dtype=torch.float32, device=h.device)
        # the fused single-pass kernel also sums the columns of the branch gradient (= bias gradient of the Linear
        # that produced the branch output): offered to that layer's backward through the colsum cache
        single_pass = cols <= 2048 and rows >= 64 and os.environ.get("QN_LN_BWD_SPLIT") is None
        fused = single_pass and ctx.needs_input_grad[0]
        cs = torch.empty(cols, dtype=h.dtype, device=h.device) if fused else None
        # fp32 main grads: dgamma / dbeta are accumulated in place by the kernel (no finalize pass, no adds by autograd)
        from . import fused_wgrad_accumulation, wgrad_done
        wp, bp = getattr(ctx, "ln_params", (None, None))
        mg_w, mg_b = getattr(wp, "main_grad", None), getattr(bp, "main_grad", None)
        in_place = (single_pass and fused_wgrad_accumulation() and mg_w is not None and mg_b is not None
                    and mg_w.dtype == torch.float32 and mg_b.dtype == torch.float32 and mg_w.numel() == cols
                    and mg_b.numel() == cols and mg_w.is_contiguous() and mg_b.is_contiguous())
        rc = _lib().qn_dropout_add_layernorm_bwd(dy2.data_ptr(), h.data_ptr(), w.data_ptr(), mean.data_ptr(),
                                                 rstd.data_ptr(), native.ptr(dh2), dres.data_ptr(), native.ptr(dx),
                                                 dg.data_ptr(), db.data_ptr(), ws.data_ptr(), nparts, rows, cols,
                                                 p, seed, off, native.ptr(cs), mg_w.data_ptr() if in_place else None,
                                                 mg_b.data_ptr() if in_place else None, native.stream())
        native.check(rc, "qn_dropout_add_layernorm_bwd")
        native.count(2)
        dres_v = dres.view(shape)
        dx_v = dx.view(shape) if dx is not None else dres_v
        if cs is not None and dx is not None:        # (p == 0: dx aliases dres, which also feeds the residual stream)
            offer_colsum(dx_v, cs)
        if in_place:
            wgrad_done(wp)
            return dx_v, dres_v, None, None, None, None
        return dx_v, dres_v, dg.to(wdtype), db.to(wdtype), None, None


def dropout_add_layer_norm(x, residual, w, b, eps, p):
    return _DropoutAddLayerNorm.apply(x, residual, w, b, eps, p)


def dropout_add_layer_norm_ok(x) -> bool:
    cols = x.shape[-1]
    return x.dtype == torch.bfloat16 and cols % 8 == 0 and cols <= 2048


# --------------------------------------------------------------------- RoPE
class _Rope(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cos, sin):
        B, H, T, d = x.shape
        assert x.stride(-1) == 1
        y = torch.empty((B, H, T, d), dtype=x.dtype, device=x.device)
        cosf, sinf = cos.float().contiguous(), sin.float().contiguous()
        native.check(_lib().qn_rope(x.data_ptr(), y.data_ptr(), cosf.data_ptr(), sinf.data_ptr(), B, H, T, d,
                                    x.stride(0), x.stride(1), x.stride(2), 0, native.stream()), "qn_rope")
        native.count()
        ctx.save_for_backward(cosf, sinf)
        return y

    @staticmethod
    def backward(ctx, dy):
        cosf, sinf = ctx.saved_tensors
        dyc = dy.contiguous()
        B, H, T, d = dyc.shape
        dx = torch.empty_like(dyc)
        native.check(_lib().qn_rope(dyc.data_ptr(), dx.data_ptr(), cosf.data_ptr(), sinf.data_ptr(), B, H, T, d,
                                    dyc.stride(0), dyc.stride(1), dyc.stride(2), 1, native.stream()), "qn_rope")
        native.count()
        return dx, None, None


def rope(x, cos, sin):
    return _Rope.apply(x, cos, sin)


# ------------------------------------------------------------ cross entropy
def _round_up(x, m):
    return (x + m - 1) // m * m


def _ce_inplace(logits_buf: torch.Tensor, labels: torch.Tensor, V: int, ignore_index: int, write_grad: bool):
    stats = torch.zeros(2, dtype=torch.float32, device=logits_buf.device)
    lab = labels.contiguous()
    if lab.dtype != torch.int64:
        lab = lab.long()
    rc = _lib().qn_cross_entropy(logits_buf.data_ptr(), lab.data_ptr(), stats.data_ptr(),
                                 stats.data_ptr() + 4, logits_buf.shape[0], V, logits_buf.stride(0),
                                 ignore_index, int(write_grad), native.stream())
    native.check(rc, "qn_cross_entropy")
    native.count()
    return stats[0], stats[1]


def _scale(x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(x)
    gf = g.reshape(1).float()
    native.check(_lib().qn_scale_bf16(x.data_ptr(), out.data_ptr(), gf.data_ptr(), x.numel(), native.stream()),
                 "qn_scale_bf16")
    native.count()
    return out


class _CrossEntropySum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        N, V = logits.shape
        Vp = _round_up(V, 8)
        buf = torch.empty((N, Vp), dtype=torch.bfloat16, device=logits.device)
        buf[:, :V].copy_(logits)
        need = ctx.needs_input_grad[0]
        loss, n = _ce_inplace(buf, labels, V, ignore_index, need)
        if need:
            ctx.save_for_backward(buf)
        ctx.V, ctx.dtype = V, logits.dtype
        ctx.mark_non_differentiable(n)
        return loss, n

    @staticmethod
    def backward(ctx, g, _gn):
        (buf,) = ctx.saved_tensors
        return (buf[:, :ctx.V].float() * g).to(ctx.dtype), None, None


def cross_entropy_sum(logits, labels, ignore_index=-100):
    return _CrossEntropySum.apply(logits, labels, ignore_index)


class _LinearCrossEntropy(torch.autograd.Function):
    """loss_sum, n_valid = CE(h · Wᵀ, labels).  Forward computes the bf16 logits with the
    tcgen05 GEMM, turns them IN PLACE into dlogits, and immediately produces the (unscaled)
    dH and dW with two more GEMMs; backward only scales them by the upstream scalar."""

    @staticmethod
    def forward(ctx, h, w, labels, ignore_index):
        N, E = h.shape
        V = w.shape[0]
        Vp = _round_up(V, 64)
        h2 = h if h.is_contiguous() else h.contiguous()
        wq = w if w.dtype == torch.bfloat16 else w.to(torch.bfloat16)
        need_h, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        logits = torch.empty((N, Vp), dtype=torch.bfloat16, device=h.device)
        gemm(h2, wq, N, V, E, out=logits[:, :V])
        loss, n = _ce_inplace(logits, labels, V, ignore_index, need_h or need_w)
        dh = dw = None
        if need_h:      # dH[N,E] = dlogits[N,V] · W[V,E]
            dh = gemm(logits, wq, N, E, V, b_mn=True)
        if need_w:      # dW[V,E] = dlogitsᵀ[V,N] · H[N,E]
            dw = gemm(logits, h2, V, E, N, a_mn=True, b_mn=True)
        del logits
        ctx.save_for_backward(dh, dw)
        ctx.wdtype = w.dtype
        ctx.mark_non_differentiable(n)
        return loss, n

    @staticmethod
    def backward(ctx, g, _gn):
        dh, dw = ctx.saved_tensors
        gh = _scale(dh, g) if dh is not None else None
        gw = _scale(dw, g) if dw is not None else None
        if gw is not None and ctx.wdtype != torch.bfloat16:
            gw = gw.to(ctx.wdtype)
        return gh, gw, None, None


def linear_cross_entropy(h, w, labels, ignore_index=-100):
    N, E = h.shape
    if E % 8 or (h.data_ptr() % 16) or (w.data_ptr() % 16) or N % 8:
        from . import reference
        return reference.linear_cross_entropy(h, w.to(h.dtype), labels, ignore_index)
    return _LinearCrossEntropy.apply(h, w, labels, ignore_index)


# ---------------------------------------------------------------- attention
class _AttnPacked(torch.autograd.Function):
    """Flash attention on the packed c_attn output: qkv [B, T, 3*H*64] -> out [B, T, H*64]."""

    @staticmethod
    def forward(ctx, qkv, H, causal, p_drop, scale):
        B, T, _ = qkv.shape
        qkv_c = qkv if qkv.is_contiguous() else qkv.contiguous()
        out = torch.empty((B, T, H * 64), dtype=torch.bfloat16, device=qkv.device)
        lse = torch.empty((B, H, T), dtype=torch.float32, device=qkv.device)
        seed = 0
        if p_drop > 0.0:
            seed, off = _PhiloxState.next(8)
            seed = (seed * 0x9E3779B97F4A7C15 + off) & 0xFFFFFFFFFFFFFFFF
        rc = _lib().qn_attn_fwd(qkv_c.data_ptr(), out.data_ptr(), lse.data_ptr(), B, H, T, float(scale), int(causal),
                                float(p_drop), seed, native.stream())
        native.check(rc, "qn_attn_fwd")
        native.count()
        ctx.save_for_backward(qkv_c, out, lse)
        ctx.args = (B, H, T, float(scale), int(causal), float(p_drop), seed)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        B, H, T, scale, causal, p_drop, seed = ctx.args
        do = dout if dout.is_contiguous() else dout.contiguous()
        dqkv = torch.empty_like(qkv)
        ws = torch.empty(B * T * H * 64 + B * H * T, dtype=torch.float32, device=qkv.device)
        rc = _lib().qn_attn_bwd(qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                ws.data_ptr(), B, H, T, scale, causal, p_drop, seed, native.stream())
        native.check(rc, "qn_attn_bwd")
        native.count(5)
        return dqkv, None, None, None, None


def attention_qkvpacked(qkv: torch.Tensor, H: int, causal: bool, p_drop: float = 0.0, scale=None):
    if scale is None:
        scale = 1.0 / 8.0   # 1/sqrt(64)
    return _AttnPacked.apply(qkv, H, causal, p_drop, scale)


# ------------------------------------------------- short-sequence attention (whole head in shared memory)
def small_attention_ok(q, k, v, dropout_p: float) -> bool:
    """True if the one-CTA-per-head kernel (``csrc/attention_small.cu``) can take these [B, H, T, d] views."""
    if not (q.is_cuda and q.dtype == k.dtype == v.dtype == torch.bfloat16 and dropout_p == 0.0 and q.dim() == 4):
        return False
    if q.shape != k.shape or q.shape != v.shape:
        return False
    T, d = q.shape[2], q.shape[3]
    if not _lib().qn_small_attn_fits(T, d):
        return False
    return all(t.stride(3) == 1 and all(int(st) % 8 == 0 for st in t.stride()[:3]) and t.data_ptr() % 16 == 0
               for t in (q, k, v))


def _strides3(*ts):
    import ctypes
    vals = []
    for t in ts:
        vals += [t.stride(0), t.stride(1), t.stride(2)]
    return (ctypes.c_longlong * len(vals))(*vals)


class _SmallAttention(torch.autograd.Function):
    """softmax(Q·Kᵀ·scale [causal]) · V with Q, K, V, scores and probabilities resident in one CTA's shared memory.
    Inputs are [B, H, T, d] VIEWS (any batch / head / token strides); the output is a [B, H, T, d] view of a
    contiguous [B, T, H·d] buffer, so the caller's ``transpose(1, 2).reshape(B, T, H·d)`` is free."""

    @staticmethod
    def forward(ctx, q, k, v, causal, scale):
        B, H, T, d = q.shape
        buf = torch.empty((B, T, H, d), dtype=torch.bfloat16, device=q.device)
        o = buf.permute(0, 2, 1, 3)
        need = any(ctx.needs_input_grad[:3])
        rc = _lib().qn_small_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), None, B, H, T, d,
                                      _strides3(q, k, v, o), float(scale), int(causal), native.stream())
        native.check(rc, "qn_small_attn_fwd")
        native.count()
        if need:
            ctx.save_for_backward(q, k, v)
        ctx.args = (B, H, T, d, float(scale), int(causal))
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v = ctx.saved_tensors
        B, H, T, d, scale, causal = ctx.args
        if do.stride(3) != 1 or any(int(st) % 8 for st in do.stride()[:3]) or do.data_ptr() % 16:
            do = do.contiguous()
        dq, dk, dv = (torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, device=t.device) for t in (q, k, v))
        rc = _lib().qn_small_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), do.data_ptr(), dq.data_ptr(),
                                      dk.data_ptr(), dv.data_ptr(), B, H, T, d, _strides3(q, k, v, do), scale, causal,
                                      native.stream())
        native.check(rc, "qn_small_attn_bwd")
        native.count()
        return dq, dk, dv, None, None


def small_attention(q, k, v, causal: bool, scale=None):
    scale = (q.shape[-1] ** -0.5) if scale is None else scale
    return _SmallAttention.apply(q, k, v, bool(causal), float(scale))
