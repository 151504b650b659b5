"""Single-GPU "loop-back" tests of the fused tensor-parallel kernels (K3 AG->GEMM, K4 GEMM->RS).

Every rank of a world of 2 (or 4) is ONE launch of the real kernel on its own CUDA stream of the SAME device, capped
to 148/world CTAs so that all "ranks" are co-resident; the peer pointers are ordinary local allocations.  The whole
protocol runs for real — copy pump, per-unit flags, TMA flag-acquire loads, staged partial slabs, lazy flag
publication, own-tiles-last reduction, CTA pairs — only the NVLink hop is missing, so the driver's 1-GPU test box
(and any 1-GPU developer box) exercises the same binary code that `tests/test_comm_gpu.py` runs across GPUs.

Oracle: plain PyTorch fp32 of the same op.
"""
import ctypes
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_MUL_AUX = 0, 1, 2, 3


def _lib():
    from quintnet_b200.ops import native, tp_fused  # noqa: F401  (declares the fused signatures)
    return native.require()


def _ptrs(ts):
    return (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def _rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


def _set_env(pairs, mode, parts=1):
    os.environ["QN_TP_2CTA"] = "1" if pairs else "0"
    os.environ["QN_TP_AG_MODE"] = mode
    os.environ["QN_TP_AG_PARTS"] = os.environ["QN_TP_RS_PARTS"] = str(parts)   # 4 = opt-in per-part publication


@pytest.fixture(autouse=True)
def _restore_env():
    keep = {k: os.environ.get(k) for k in ("QN_TP_2CTA", "QN_TP_AG_MODE", "QN_TP_AG_PARTS", "QN_TP_RS_PARTS")}
    yield
    for k, v in keep.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _run_ag(world, M, N, K, b_mn, epi, pairs, mode, launches=3):
    lib = _lib()
    rows = M // world
    xs = [_rnd(rows, K) for _ in range(world)]
    ws = [_rnd(K, N, scale=K ** -0.5) if b_mn else _rnd(N, K, scale=K ** -0.5) for _ in range(world)]
    bias = [_rnd(N) if epi in (EPI_BIAS, EPI_BIAS_GELU) else None for _ in range(world)]
    aux = [_rnd(M, N) if epi == EPI_MUL_AUX else None for _ in range(world)]
    gather = [[torch.zeros(M * K, dtype=torch.bfloat16, device="cuda") for _ in range(world)] for _ in range(2)]
    flags = [torch.zeros(1 << 18, dtype=torch.int32, device="cuda") for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    cap = (148 // world) & ~1
    _set_env(pairs, mode, 4 if M >= 16384 else 1)             # bench-size cases also cover the per-part protocol
    torch.cuda.synchronize()
    for seq in range(1, launches + 1):                       # several launches: double-buffered workspaces + seq flags
        ys = [torch.empty(M, N, dtype=torch.bfloat16, device="cuda") for _ in range(world)]
        zs = [torch.empty(M, N, dtype=torch.bfloat16, device="cuda") if epi == EPI_BIAS_GELU else aux[r] for r in range(world)]
        par = seq & 1
        # odd launches also ask for the private gathered-operand tensor (forward mode), even ones use the buffer view
        xfull = [torch.zeros(M, K, dtype=torch.bfloat16, device="cuda") if (mode == "pump" and seq % 2 == 1) else None
                 for _ in range(world)]
        torch.cuda.synchronize()                             # the zero-fills above ran on the default stream
        for r in range(world):
            rc = lib.qn_gemm_ag_bf16(xs[r].data_ptr(), _ptrs(gather[par]), _ptrs(flags), r, world, seq, ws[r].data_ptr(),
                                     ys[r].data_ptr(), M, N, K, ws[r].stride(0), N, int(b_mn), epi,
                                     bias[r].data_ptr() if bias[r] is not None else None,
                                     zs[r].data_ptr() if zs[r] is not None else None, N if zs[r] is not None else 0,
                                     0, cap, xfull[r].data_ptr() if xfull[r] is not None else None,
                                     streams[r].cuda_stream)
            assert rc == 0, rc
        torch.cuda.synchronize()
        xf = torch.cat(xs).float()
        for r in range(world):
            w = ws[r].float()
            ref = xf @ (w if b_mn else w.t())
            if bias[r] is not None:
                ref = ref + bias[r].float()
            if epi == EPI_BIAS_GELU:
                z = ref.clone().requires_grad_(True)
                act = torch.nn.functional.gelu(z)
                act.sum().backward()
                torch.testing.assert_close(zs[r].float(), z.grad, atol=3e-2, rtol=3e-2)
                ref = act.detach()
            if epi == EPI_MUL_AUX:
                ref = ref * aux[r].float()
            torch.testing.assert_close(ys[r].float(), ref, atol=6e-2, rtol=3e-2)
            # the kernel leaves the complete gathered operand in the private tensor / in every rank's gather buffer
            if xfull[r] is not None:
                torch.testing.assert_close(xfull[r].float(), xf, atol=0, rtol=0)
            elif mode != "ce":
                torch.testing.assert_close(gather[par][r].view(M, K).float(), xf, atol=0, rtol=0)
        xs = [_rnd(rows, K) for _ in range(world)]           # new data next launch (stale buffers must not pass)


def _run_rs(world, M, N, K, b_mn, with_bias, pairs, launches=3):
    lib = _lib()
    rows = M // world
    as_ = [_rnd(M, K) for _ in range(world)]
    ws = [_rnd(K, N, scale=(K * world) ** -0.5) if b_mn else _rnd(N, K, scale=(K * world) ** -0.5) for _ in range(world)]
    bias = _rnd(N) if with_bias else None
    nbytes = M * ((N + 63) // 64) * 128
    stage = [[torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(world)] for _ in range(2)]
    flags = [torch.zeros(1 << 18, dtype=torch.int32, device="cuda") for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    cap = (148 // world) & ~1
    _set_env(pairs, "pump", 4 if M >= 16384 else 1)
    torch.cuda.synchronize()
    for seq in range(1, launches + 1):
        ys = [torch.empty(rows, N, dtype=torch.bfloat16, device="cuda") for _ in range(world)]
        par = seq & 1
        for r in range(world):
            rc = lib.qn_gemm_rs_bf16(as_[r].data_ptr(), ws[r].data_ptr(), ys[r].data_ptr(), _ptrs(stage[par]), _ptrs(flags),
                                     r, world, seq, M, N, K, as_[r].stride(0), ws[r].stride(0), N, int(b_mn),
                                     bias.data_ptr() if (bias is not None and r == 0) else None, 0, cap,
                                     streams[r].cuda_stream)   # the bias lives on tp-rank 0 only
            assert rc == 0, rc
        torch.cuda.synchronize()
        total = sum(as_[r].float() @ (ws[r].float() if b_mn else ws[r].float().t()) for r in range(world))
        if bias is not None:
            total = total + bias.float()
        for r in range(world):
            torch.testing.assert_close(ys[r].float(), total[r * rows:(r + 1) * rows], atol=8e-2, rtol=3e-2)
        as_ = [_rnd(M, K) for _ in range(world)]


@pytest.mark.timeout(180)
@pytest.mark.parametrize("pairs", [True, False], ids=["cta-pairs", "single-cta"])
@pytest.mark.parametrize("mode", ["pump", "lsu"])
@pytest.mark.parametrize("world,M,N,K,b_mn,epi", [
    (2, 4096, 1536, 1024, False, EPI_BIAS),          # c_attn forward shape (per-rank N), GPT-2 medium
    (2, 4096, 2048, 1024, False, EPI_BIAS_GELU),     # fc forward + GELU (+ gelu' second output)
    (2, 2048, 512, 1024, True, EPI_NONE),            # c_proj dgrad (B MN-major)
    (2, 2048, 1024, 768, True, EPI_MUL_AUX),         # mlp-proj dgrad ⊙ gelu', K = 768 (12 KB pump units)
    (4, 4096, 512, 1280, False, EPI_NONE),           # world 4, K = 1280 (20 KB units -> two pump pieces)
    (2, 16384, 512, 1024, True, EPI_NONE),           # bench-size shard (8192 rows): four separately published parts
])
def test_ag_gemm_loopback(world, M, N, K, b_mn, epi, pairs, mode):
    torch.manual_seed(0)
    _run_ag(world, M, N, K, b_mn, epi, pairs, mode)


@pytest.mark.timeout(180)
@pytest.mark.parametrize("pairs", [True, False], ids=["cta-pairs", "single-cta"])
@pytest.mark.parametrize("world,M,N,K,b_mn,with_bias", [
    (2, 4096, 1024, 512, False, True),               # c_proj forward
    (2, 4096, 1024, 2048, False, True),              # mlp-proj forward
    (2, 2048, 1024, 1536, True, False),              # c_attn dgrad (B MN-major)
    (2, 2048, 1000, 512, False, True),               # N not a multiple of 64 (clipped last chunk)
    (4, 4096, 768, 256, False, True),                # world 4: three staged partials per slab
    (2, 16384, 1024, 512, False, True),              # bench-size shard: per-part publication (4 parts per owner)
    (4, 16384, 512, 256, False, True),               # world 4 with parts: 12 (owner, part) keys per rank
])
def test_gemm_rs_loopback(world, M, N, K, b_mn, with_bias, pairs):
    torch.manual_seed(0)
    _run_rs(world, M, N, K, b_mn, with_bias, pairs)
