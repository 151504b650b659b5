"""sm_100a kernel numerics vs plain-PyTorch fp32 references (run on the B200 box)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from quintnet_b200.ops import native
    return native.require()


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 128), (384, 768, 1024), (16384 // 8, 3072, 1024),
                                   (400, 64, 64), (130, 264, 72), (1000, 6288, 256)])
@pytest.mark.parametrize("block_n", [0, 128, 256])
def test_gemm_tn(M, N, K, block_n):
    from quintnet_b200.ops import cuda_ops
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    c = cuda_ops.gemm(a, b, M, N, K, block_n=block_n)
    ref = a.float() @ b.float().t()
    assert _rel(c, ref) < 1e-2, _rel(c, ref)


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (384, 1024, 520), (1024, 768, 2048), (136, 200, 72)])
def test_gemm_layouts(a_mn, b_mn, M, N, K):
    from quintnet_b200.ops import cuda_ops
    torch.manual_seed(1)
    A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    B = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    c = cuda_ops.gemm(a, b, M, N, K, a_mn=a_mn, b_mn=b_mn)
    ref = A.float() @ B.float().t()
    assert _rel(c, ref) < 1e-2, _rel(c, ref)


def test_gemm_epilogues():
    from quintnet_b200.ops import cuda_ops
    torch.manual_seed(2)
    M, N, K = 512, 1024, 256
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16) * 0.5
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.1
    bias = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    ref = a.float() @ b.float().t() + bias.float()
    c = cuda_ops.gemm(a, b, M, N, K, epilogue=cuda_ops.EPI_BIAS, bias=bias)
    assert _rel(c, ref) < 1e-2
    z = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    g = cuda_ops.gemm(a, b, M, N, K, epilogue=cuda_ops.EPI_BIAS_GELU, bias=bias, z=z)
    dg = 0.5 * (1 + torch.erf(ref / math.sqrt(2))) + ref * torch.exp(-0.5 * ref * ref) / math.sqrt(2 * math.pi)
    assert _rel(z, dg) < 1e-2, "second output = gelu'(pre-activation)"
    assert _rel(g, F.gelu(ref)) < 1e-2
    aux = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    gb = cuda_ops.gemm(a, b, M, N, K, epilogue=cuda_ops.EPI_MUL_AUX, z=aux)
    assert _rel(gb, (a.float() @ b.float().t()) * aux.float()) < 1.5e-2
    for (m2, n2) in [(130, 200), (1000, 328)]:          # ragged tiles through the TMA aux path
        a2 = torch.randn(m2, K, device="cuda", dtype=torch.bfloat16)
        b2 = torch.randn(n2, K, device="cuda", dtype=torch.bfloat16) * 0.1
        x2 = torch.randn(m2, n2, device="cuda", dtype=torch.bfloat16)
        o2 = cuda_ops.gemm(a2, b2, m2, n2, K, epilogue=cuda_ops.EPI_MUL_AUX, z=x2)
        assert _rel(o2, (a2.float() @ b2.float().t()) * x2.float()) < 1.5e-2


def test_linear_autograd():
    from quintnet_b200 import ops
    torch.manual_seed(3)
    for act in (None, "gelu"):
        x = torch.randn(4, 96, 256, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        w = (torch.randn(512, 256, device="cuda", dtype=torch.bfloat16) * 0.05).requires_grad_(True)
        b = torch.randn(512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        y = ops.linear(x, w, b, act)
        dy = torch.randn_like(y)
        y.backward(dy)
        xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, b))
        yr = F.linear(xr, wr, br)
        if act == "gelu":
            yr = F.gelu(yr)
        yr.backward(dy.float())
        assert _rel(y, yr) < 1e-2
        assert _rel(x.grad, xr.grad) < 2e-2, (act, _rel(x.grad, xr.grad))
        assert _rel(w.grad, wr.grad) < 2e-2
        assert _rel(b.grad, br.grad) < 2e-2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("rows,cols", [(64, 64), (1000, 768), (4096, 1024), (333, 1600)])
def test_layernorm(dtype, rows, cols):
    from quintnet_b200 import ops
    torch.manual_seed(4)
    x = torch.randn(rows, cols, device="cuda", dtype=dtype, requires_grad=True)
    w = torch.randn(cols, device="cuda", dtype=dtype, requires_grad=True)
    b = torch.randn(cols, device="cuda", dtype=dtype, requires_grad=True)
    y = ops.layer_norm(x, w, b, 1e-5)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    yr = F.layer_norm(xr, (cols,), wr, br, 1e-5)
    yr.backward(dy.float())
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4
    assert _rel(y, yr) < tol
    assert _rel(x.grad, xr.grad) < tol
    assert _rel(w.grad, wr.grad) < tol and _rel(b.grad, br.grad) < tol


def test_add_layernorm():
    from quintnet_b200 import ops
    torch.manual_seed(5)
    x = torch.randn(512, 768, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    r = torch.randn(512, 768, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    w = torch.ones(768, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    b = torch.zeros(768, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y, s = ops.add_layer_norm(x, r, w, b)
    (y.float().pow(2).sum() + s.float().sum()).backward()
    xr, rr = x.detach().float().requires_grad_(True), r.detach().float().requires_grad_(True)
    sr = xr + rr
    yr = F.layer_norm(sr, (768,))
    (yr.pow(2).sum() + sr.sum()).backward()
    assert _rel(y, yr) < 2e-2 and _rel(s, sr) < 1e-2
    assert _rel(x.grad, xr.grad) < 3e-2 and _rel(r.grad, rr.grad) < 3e-2


def test_gelu_dropout_rope():
    from quintnet_b200 import ops
    from quintnet_b200.ops import cuda_ops, reference
    torch.manual_seed(6)
    x = torch.randn(1024, 512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y = ops.gelu(x)
    y.backward(torch.ones_like(y))
    xr = x.detach().float().requires_grad_(True)
    F.gelu(xr).sum().backward()
    assert _rel(y, F.gelu(xr)) < 1e-2 and _rel(x.grad, xr.grad) < 1e-2
    # dropout: keep fraction, scaling, mask consistency between fwd and bwd
    cuda_ops.manual_seed(123)
    x = torch.ones(1 << 20, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    res = torch.full_like(x, 2.0)
    y = ops.dropout_add(x, res, 0.1, True)
    kept = (y.float() > 2.5)
    assert abs(kept.float().mean().item() - 0.9) < 5e-3
    assert torch.allclose(y.float()[kept], torch.full_like(y.float()[kept], 2.0 + 1 / 0.9), atol=3e-2)
    y.backward(torch.ones_like(y))
    assert torch.equal(x.grad.float() > 0.5, kept)
    # rope
    q = torch.randn(2, 4, 64, 64, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    cos, sin = ops.rope_tables(64, 64, device="cuda")
    o = ops.rope(q, cos, sin)
    o.backward(torch.ones_like(o))
    qr = q.detach().float().requires_grad_(True)
    orf = reference.rope(qr, cos, sin)
    orf.sum().backward()
    assert _rel(o, orf) < 1e-2 and _rel(q.grad, qr.grad) < 1e-2


def test_fused_adamw_matches_torch():
    from quintnet_b200.optim import FusedAdamW
    torch.manual_seed(7)
    ref = torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.Linear(128, 32)).cuda()
    mine = torch.nn.Sequential(torch.nn.Linear(64, 128), torch.nn.Linear(128, 32)).cuda()
    mine.load_state_dict(ref.state_dict())
    o_ref = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.01)
    o_mine = FusedAdamW(mine.parameters(), lr=1e-2, weight_decay=0.01)
    for _ in range(5):
        x = torch.randn(16, 64, device="cuda")
        for m, o in ((ref, o_ref), (mine, o_mine)):
            o.zero_grad()
            m(x).pow(2).mean().backward()
            o.step()
    for p, q in zip(ref.parameters(), mine.parameters()):
        assert torch.allclose(p, q, atol=1e-5), (p - q).abs().max()


def test_linear_cross_entropy():
    from quintnet_b200 import ops
    torch.manual_seed(8)
    N, E, V = 1024, 256, 50257 // 16
    h = (torch.randn(N, E, device="cuda", dtype=torch.bfloat16) * 0.5).requires_grad_(True)
    w = (torch.randn(V, E, device="cuda", dtype=torch.bfloat16) * 0.05).requires_grad_(True)
    labels = torch.randint(0, V, (N,), device="cuda")
    labels[::7] = -100
    ls, n = ops.linear_cross_entropy(h, w, labels)
    (ls / n).backward()
    hr, wr = h.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    lr_ = F.cross_entropy(hr @ wr.t(), labels, ignore_index=-100, reduction="sum")
    nr = (labels != -100).sum()
    (lr_ / nr).backward()
    assert abs(ls.item() - lr_.item()) / lr_.item() < 5e-3 and int(n.item()) == int(nr.item())
    assert _rel(h.grad, hr.grad) < 3e-2, _rel(h.grad, hr.grad)
    assert _rel(w.grad, wr.grad) < 3e-2, _rel(w.grad, wr.grad)


def test_gpt2_gpu_matches_cpu_reference():
    """Tiny GPT-2 in bf16 through the native ops vs the same weights in fp32 on CPU."""
    from quintnet_b200.models.gpt2 import GPT2Config, GPT2Model
    cfg = GPT2Config(vocab_size=512, n_positions=64, n_embd=128, n_layer=2, n_head=4,
                     resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    ref = GPT2Model(cfg, seed=3)
    gpu = GPT2Model(cfg, device=torch.device("cuda"), dtype=torch.bfloat16, seed=3)
    gpu.load_state_dict(ref.state_dict())      # CPU and CUDA generators differ: copy the weights
    ids = torch.randint(0, 512, (8, 64))
    lr_, nr = ref(ids, labels=ids)
    lg, ng = gpu(ids.cuda(), labels=ids.cuda())
    assert abs(lg.item() / ng.item() - lr_.item() / nr.item()) < 1e-2
    (lr_ / nr).backward()
    (lg / ng).backward()
    g1 = ref.blocks[0].mlp.c_fc.proj.weight.grad
    g2 = gpu.blocks[0].mlp.c_fc.proj.weight.grad
    assert _rel(g2.cpu(), g1) < 8e-2, _rel(g2.cpu(), g1)


@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("B,H,T", [(2, 4, 128), (2, 3, 512), (1, 2, 384), (2, 2, 200)])
def test_attention_qkvpacked(causal, B, H, T):
    from quintnet_b200 import ops
    from quintnet_b200.ops import reference
    torch.manual_seed(9)
    qkv = (torch.randn(B, T, 3 * H * 64, device="cuda", dtype=torch.bfloat16)).requires_grad_(True)
    out = ops.attention_qkvpacked(qkv, H, causal=causal, dropout_p=0.0)
    dout = torch.randn_like(out)
    out.backward(dout)
    qr = qkv.detach().float().requires_grad_(True)
    q, k, v = qr.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    oref = reference.attention(q, k, v, causal, 0.0, None, False).transpose(1, 2).reshape(B, T, H * 64)
    oref.backward(dout.float())
    assert _rel(out, oref) < 2e-2, _rel(out, oref)
    g, gr = qkv.grad.view(B, T, 3, H, 64), qr.grad.view(B, T, 3, H, 64)
    for i, name in enumerate("qkv"):
        assert _rel(g[:, :, i], gr[:, :, i]) < 4e-2, (name, _rel(g[:, :, i], gr[:, :, i]))


def test_attention_dropout_statistics():
    from quintnet_b200 import ops
    from quintnet_b200.ops import cuda_ops
    cuda_ops.manual_seed(5)
    B, H, T = 2, 4, 256
    qkv = torch.zeros(B, T, 3 * H * 64, device="cuda", dtype=torch.bfloat16)
    qkv.view(B, T, 3, H, 64)[:, :, 2] = 1.0            # V = 1, uniform attention -> out = kept fraction / (1-p)
    qkv.requires_grad_(True)
    out = ops.attention_qkvpacked(qkv, H, causal=False, dropout_p=0.25)
    m = out.float().mean().item()
    assert abs(m - 1.0) < 2e-2, m                       # E[keep/(1-p)] = 1
    assert out.float().std().item() > 1e-3              # masks differ per row
    out.sum().backward()                                # backward regenerates the same mask (finite grads)
    assert torch.isfinite(qkv.grad.float()).all()


def test_fused_mlp():
    from quintnet_b200 import ops
    torch.manual_seed(10)
    x = torch.randn(512, 256, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    w1 = (torch.randn(1024, 256, device="cuda", dtype=torch.bfloat16) * 0.05).requires_grad_(True)
    b1 = (torch.randn(1024, device="cuda", dtype=torch.bfloat16) * 0.1).requires_grad_(True)
    w2 = (torch.randn(256, 1024, device="cuda", dtype=torch.bfloat16) * 0.05).requires_grad_(True)
    b2 = (torch.randn(256, device="cuda", dtype=torch.bfloat16) * 0.1).requires_grad_(True)
    y = ops.mlp(x, w1, b1, w2, b2)
    dy = torch.randn_like(y)
    y.backward(dy)
    ref = [t.detach().float().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    yr = F.linear(F.gelu(F.linear(ref[0], ref[1], ref[2])), ref[3], ref[4])
    yr.backward(dy.float())
    assert _rel(y, yr) < 2e-2
    for t, r, n in zip((x, w1, b1, w2, b2), ref, "x w1 b1 w2 b2".split()):
        assert _rel(t.grad, r.grad) < 3e-2, (n, _rel(t.grad, r.grad))


@pytest.mark.parametrize("M,N,K,splits", [(1024, 1024, 16384, 0), (1024, 512, 16384, 4), (256, 256, 1000, 3),
                                          (1536, 1024, 8192, 0)])
def test_gemm_split_k(M, N, K, splits):
    from quintnet_b200.ops import cuda_ops
    torch.manual_seed(11)
    A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    B = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    ref = A.float() @ B.float().t()
    c = cuda_ops.gemm(A.t().contiguous(), B.t().contiguous(), M, N, K, a_mn=True, b_mn=True, split_k=splits)
    assert _rel(c, ref) < 1e-2, _rel(c, ref)
    c2 = cuda_ops.gemm(A, B, M, N, K, split_k=splits if splits else 0)
    assert _rel(c2, ref) < 1e-2


def test_dropout_add_layernorm_fused():
    from quintnet_b200 import ops
    from quintnet_b200.ops import cuda_ops
    torch.manual_seed(12)
    for p in (0.0, 0.1):
        x = torch.randn(1024, 1024, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        r = torch.randn(1024, 1024, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        w = torch.randn(1024, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        b = torch.randn(1024, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        cuda_ops.manual_seed(77)
        y, h = ops.dropout_add_layer_norm(x, r, w, b, 1e-5, p, True)
        gy, gh = torch.randn_like(y), torch.randn_like(h)
        torch.autograd.backward([y, h], [gy, gh])
        # reference with the SAME mask: recover it from h - r  (kept elements have |h - r| > 0 almost surely)
        keep = ((h.detach().float() - r.detach().float()).abs() > 1e-6) if p > 0 else torch.ones_like(h, dtype=torch.bool)
        if p > 0:
            assert abs(keep.float().mean().item() - (1 - p)) < 1e-2
        xr, rr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, r, w, b))
        hr = xr * keep.float() / (1 - p) + rr
        yr = F.layer_norm(hr, (1024,), wr, br, 1e-5)
        torch.autograd.backward([yr, hr], [gy.float(), gh.float()])
        assert _rel(h, hr) < 1e-2 and _rel(y, yr) < 2e-2
        assert _rel(x.grad, xr.grad) < 6e-2, (p, _rel(x.grad, xr.grad))
        assert _rel(r.grad, rr.grad) < 3e-2
        assert _rel(w.grad, wr.grad) < 3e-2 and _rel(b.grad, br.grad) < 3e-2


@pytest.mark.parametrize("M,N,K", [(256, 512, 256), (1000, 768, 1024), (4096, 3072, 1024)])
def test_fp8_quantize_and_gemm(M, N, K):
    """e4m3 row-wise quantiser and the kind::f8f6f4 GEMM vs an fp32 emulation of the same quantisation."""
    from quintnet_b200.ops import cuda_ops, reference
    torch.manual_seed(21)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    a[:, :7] *= 30.0                       # outliers: the per-row scale has to absorb them
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    qa, sa = cuda_ops.quantize_rowwise(a)
    qw, sw = cuda_ops.quantize_rowwise(w)
    deq = qa.view(torch.float8_e4m3fn).float() * sa[:, None]
    emu = reference.fake_quant_rowwise_e4m3(a.float())
    assert _rel(deq, emu) < 5e-3          # x * (448 / amax) vs x / (amax / 448): one-ulp ties round differently
    assert torch.allclose(sa, a.float().abs().amax(dim=1) / 448.0, rtol=1e-3)
    y = cuda_ops.gemm_fp8(qa, sa, qw, sw, M, N, K)
    ref = emu @ reference.fake_quant_rowwise_e4m3(w.float()).t()
    assert _rel(y, ref) < 1e-2, _rel(y, ref)
    assert _rel(y, a.float() @ w.float().t()) < 6e-2          # and it is a usable approximation of the real product


@pytest.mark.parametrize("act", [None, "gelu"])
def test_fp8_linear_autograd(act):
    from quintnet_b200 import ops
    from quintnet_b200.ops import reference
    torch.manual_seed(22)
    x = torch.randn(8, 128, 1024, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    w = (torch.randn(2048, 1024, device="cuda") * 0.03).to(torch.bfloat16).requires_grad_(True)
    b = torch.randn(2048, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y = ops.linear(x, w, b, act, fp8=True)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    yr = reference.linear_fp8(xr, wr, br, act, recipe=ops.fp8_recipe() if x.shape[-1] % 128 == 0 else "rowwise")
    yr.backward(dy.float())
    assert _rel(y, yr) < 2e-2, _rel(y, yr)
    assert _rel(x.grad, xr.grad) < 5e-2 and _rel(w.grad, wr.grad) < 5e-2 and _rel(b.grad, br.grad) < 2e-2


def test_activation_checkpointing_replays_native_dropout():
    """Recompute-in-backward must regenerate the SAME native dropout masks (counter rewind), so gradients match."""
    from quintnet_b200.models.gpt2 import GPT2Config, GPT2Model
    from quintnet_b200.ops import cuda_ops
    base = dict(vocab_size=512, n_positions=128, n_embd=256, n_layer=2, n_head=4, embd_pdrop=0.1, attn_pdrop=0.1,
                resid_pdrop=0.1)
    ids = torch.randint(0, 512, (8, 128), device="cuda")
    outs = []
    for ckpt in (False, True):
        model = GPT2Model(GPT2Config(**dict(base, activation_checkpointing=ckpt)), device=torch.device("cuda"),
                          dtype=torch.bfloat16, seed=0).train()
        cuda_ops.manual_seed(99)
        torch.manual_seed(99)
        ls, n = model(ids, labels=ids)
        (ls / n).backward()
        outs.append((float((ls / n).detach()), torch.cat([p.grad.float().flatten() for p in model.parameters()])))
    assert abs(outs[0][0] - outs[1][0]) < 1e-3
    assert _rel(outs[1][1], outs[0][1]) < 2e-2      # attention bwd accumulates dQ with atomics: not bitwise


@pytest.mark.parametrize("N,K,M", [(1024, 1024, 4096), (4096, 1024, 16384), (768, 3072, 2048)])
def test_wgrad_accumulate_epilogue(N, K, M):
    """dW += dYᵀ·X inside the GEMM epilogue (ACCUM kernels, CTA pairs / single CTA / split-K) vs fp32."""
    from quintnet_b200 import ops
    from quintnet_b200.ops import cuda_ops
    torch.manual_seed(31)
    dy = (torch.randn(M, N, device="cuda") * 0.1).to(torch.bfloat16)
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.nn.Parameter(torch.zeros(N, K, device="cuda", dtype=torch.bfloat16))
    w.grad = (torch.randn(N, K, device="cuda") * 3.0).to(torch.bfloat16)
    ref = w.grad.float() + dy.float().t() @ x.float()
    ops.set_fused_wgrad_accumulation(True)
    try:
        out = cuda_ops.wgrad(w, dy, x, N, K, M)
    finally:
        ops.set_fused_wgrad_accumulation(False)
    assert out is None
    assert _rel(w.grad, ref) < 1e-2, _rel(w.grad, ref)


@pytest.mark.parametrize("N,K,M", [(1024, 1024, 4096), (4096, 1024, 16384), (768, 3072, 2048), (3072, 1024, 512)])
def test_wgrad_accumulate_fp32_main_grad(N, K, M):
    """G (fp32 main grad) += dYᵀ·X through the red.add epilogue (CTA pairs / single CTA / split-K), twice in a row:
    the second call must ADD to the first (micro-batch accumulation), with fp32 accuracy of the accumulation."""
    from quintnet_b200 import ops
    from quintnet_b200.ops import cuda_ops
    torch.manual_seed(33)
    w = torch.nn.Parameter(torch.zeros(N, K, device="cuda", dtype=torch.bfloat16))
    w.main_grad = torch.randn(N, K, device="cuda") * 3.0
    ref = w.main_grad.clone()
    ops.set_fused_wgrad_accumulation(True)
    try:
        for _ in range(2):
            dy = (torch.randn(M, N, device="cuda") * 0.1).to(torch.bfloat16)
            x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            ref += dy.float().t() @ x.float()
            assert cuda_ops.wgrad(w, dy, x, N, K, M) is None
    finally:
        ops.set_fused_wgrad_accumulation(False)
    assert w.grad is None
    assert _rel(w.main_grad, ref) < 2e-5, _rel(w.main_grad, ref)       # fp32 accumulation of exact bf16 products


@pytest.mark.parametrize("B,H,T,d,causal", [(8, 4, 50, 16, False), (8, 2, 50, 16, False), (3, 4, 64, 32, True),
                                            (2, 3, 17, 64, False), (4, 4, 128, 16, True)])
def test_small_attention_fwd_bwd(B, H, T, d, causal):
    """Whole-head shared-memory attention (ViT path) vs the fp32 reference, on the strided [B, H, T, d] views the
    model passes (projection outputs viewed as heads)."""
    from quintnet_b200.ops import cuda_ops, reference
    torch.manual_seed(41)
    mk = lambda: torch.randn(B, T, H * d, device="cuda").to(torch.bfloat16).requires_grad_(True)
    xq, xk, xv = mk(), mk(), mk()
    view = lambda x: x.view(B, T, H, d).transpose(1, 2)
    q, k, v = view(xq), view(xk), view(xv)
    assert cuda_ops.small_attention_ok(q, k, v, 0.0)
    o = cuda_ops.small_attention(q, k, v, causal)
    out = o.transpose(1, 2).reshape(B, T, H * d)
    assert out.data_ptr() == o.data_ptr()                    # the head concat is a free view
    g = torch.randn_like(out)
    out.backward(g)
    qf, kf, vf = (x.detach().float().requires_grad_(True) for x in (xq, xk, xv))
    ref = reference.attention(view(qf), view(kf), view(vf), causal).transpose(1, 2).reshape(B, T, H * d)
    ref.backward(g.float())
    assert _rel(out, ref) < 1e-2, _rel(out, ref)
    for a, b in ((xq, qf), (xk, kf), (xv, vf)):
        assert _rel(a.grad, b.grad) < 2e-2, _rel(a.grad, b.grad)


def test_vit_block_uses_no_library_attention():
    """The ViT forward/backward must not launch torch SDPA: ops.attention raises instead of falling back."""
    from quintnet_b200.models import Model
    torch.manual_seed(0)
    m = Model(28, 4, 64, 1, 4, 2).to("cuda", torch.bfloat16)
    x = torch.randn(8, 1, 28, 28, device="cuda", dtype=torch.bfloat16)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m(x).float().sum().backward()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    assert any("small_attn_fwd" in n for n in names) and any("small_attn_bwd" in n for n in names), names
    assert not any(("fmha" in n.lower() or "flash" in n.lower() or "sdpa" in n.lower()) for n in names), names


def test_layernorm_bwd_offers_bias_gradient():
    """The fused LayerNorm backward also produces the column sums of the branch gradient (= bias gradient of the Linear
    that fed the branch); the next backward op picks them up instead of launching a column-sum pass."""
    from quintnet_b200.ops import cuda_ops
    torch.manual_seed(7)
    rows, cols = 4096, 1024
    x0 = torch.randn(rows, cols, device="cuda").to(torch.bfloat16).requires_grad_(True)
    r = torch.randn(rows, cols, device="cuda").to(torch.bfloat16).requires_grad_(True)
    g = torch.randn(cols, device="cuda").to(torch.bfloat16).requires_grad_(True)
    b = torch.randn(cols, device="cuda").to(torch.bfloat16).requires_grad_(True)
    x = x0 * 1.0                                       # non-leaf: its gradient is the very tensor the LN backward returned
    seen = []
    x.register_hook(lambda gr: seen.append(gr))
    y, h = cuda_ops.dropout_add_layer_norm(x, r, g, b, 1e-5, 0.1)
    cuda_ops._COLSUM_CACHE.clear()
    (y.float().pow(2).sum() + h.float().sum()).backward()
    gx = seen[0]
    ent = cuda_ops._COLSUM_CACHE.get(gx.data_ptr())
    assert ent is not None and ent[0]() is not None
    ref = gx.float().sum(0)
    got = cuda_ops.colsum(gx.reshape(rows, cols))      # served from the cache (no kernel): same values as a fresh sum
    assert gx.data_ptr() not in cuda_ops._COLSUM_CACHE
    assert _rel(got, ref) < 1e-2, _rel(got, ref)
    fresh = cuda_ops.colsum(gx.reshape(rows, cols))    # cache is empty now: the stand-alone kernel
    assert _rel(fresh, ref) < 1e-2


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (256, 384, 512), (1000, 264, 1024), (16384 // 4, 4096, 1024)])
def test_mxfp8_block_scaled_gemm(M, N, K):
    """Block-scaled FP8 GEMM (tcgen05.mma.kind::mxf8f6f4.block_scale, UE8M0 scales through tcgen05.cp) vs an fp32
    emulation of the SAME quantisation: the quantiser must be bit-exact, the GEMM exact up to fp32 accumulation order."""
    from quintnet_b200.ops import cuda_ops, reference
    torch.manual_seed(51)
    # magnitudes spread over many binades so that different 32-element blocks really get different exponents
    a = (torch.randn(M, K, device="cuda") * torch.exp2(torch.randint(-6, 7, (M, K // 32), device="cuda").float()).repeat_interleave(32, 1)).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * torch.exp2(torch.randint(-4, 5, (N, K // 32), device="cuda").float()).repeat_interleave(32, 1)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda").to(torch.bfloat16) if N % 8 == 0 else None
    a8, sfa = cuda_ops.quantize_mx(a)
    b8, sfb = cuda_ops.quantize_mx(b)
    # quantiser: element bytes and block exponents equal the reference emulation
    ea, sa = reference.mx_scales_e8m0(a)
    qa_ref = (a.float().view(M, K // 32, 32) / sa.unsqueeze(-1)).clamp(-448, 448).to(torch.float8_e4m3fn).view(M, K)
    assert torch.equal(a8.view(torch.float8_e4m3fn).float(), qa_ref.float())
    rows = torch.arange(M, device="cuda")
    for j in range(0, K // 32, max(K // 32 // 3, 1)):
        kb, jj = j // 4, j % 4
        off = ((rows // 128) * (K // 128) + kb) * 512 + (rows % 32) * 16 + ((rows % 128) // 32) * 4 + jj
        assert torch.equal(sfa[off], ea[:, j])
    c = cuda_ops.gemm_mxfp8(a8, sfa, b8, sfb, M, N, K, bias)
    ref = reference.fake_quant_mxfp8(a.float()) @ reference.fake_quant_mxfp8(b.float()).t()
    if bias is not None:
        ref = ref + bias.float()
    assert _rel(c, ref) < 6e-3, _rel(c, ref)               # bf16 output rounding only


def test_mxfp8_linear_autograd_and_recipe():
    from quintnet_b200 import ops
    from quintnet_b200.ops import reference
    torch.manual_seed(52)
    x = torch.randn(4, 64, 256, device="cuda").to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(512, 256, device="cuda") * 0.05).to(torch.bfloat16).requires_grad_(True)
    b = torch.randn(512, device="cuda").to(torch.bfloat16).requires_grad_(True)
    assert ops.fp8_recipe() == "mxfp8"
    y = ops.linear(x, w, b, "gelu", fp8=True)
    y.float().pow(2).mean().backward()
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    yr = reference.linear_fp8(xr, wr, br, "gelu", recipe="mxfp8")
    yr.pow(2).mean().backward()
    assert _rel(y, yr) < 2e-2
    assert _rel(x.grad, xr.grad) < 5e-2 and _rel(w.grad, wr.grad) < 5e-2 and _rel(b.grad, br.grad) < 5e-2


def test_bias_and_layernorm_grads_accumulate_into_main_grad():
    """fp32 main grads: the bias column-sum kernel and the fused LayerNorm backward accumulate their parameter gradients
    IN PLACE (no finalize pass, nothing for autograd to add), twice in a row = micro-batch accumulation."""
    from quintnet_b200 import ops
    from quintnet_b200.ops import cuda_ops
    torch.manual_seed(61)
    rows, cols = 2048, 1024
    mk = lambda *s: torch.nn.Parameter(torch.randn(*s, device="cuda").to(torch.bfloat16))
    g, b, bias = mk(cols), mk(cols), mk(cols)
    for p_ in (g, b, bias):
        p_.main_grad = torch.zeros(cols, device="cuda")
    ref_g = torch.zeros(cols, device="cuda"); ref_b = torch.zeros(cols, device="cuda"); ref_bias = torch.zeros(cols, device="cuda")
    ops.set_fused_wgrad_accumulation(True)
    try:
        for it in range(2):
            x = torch.randn(rows, cols, device="cuda").to(torch.bfloat16).requires_grad_(True)
            r = torch.randn(rows, cols, device="cuda").to(torch.bfloat16).requires_grad_(True)
            y, h = cuda_ops.dropout_add_layer_norm(x, r, g, b, 1e-5, 0.0)
            dy = torch.randn_like(y)
            torch.autograd.backward([y, h], [dy, torch.zeros_like(h)])
            assert g.grad is None and b.grad is None
            hf = (x.float() + r.float()).detach()
            xhat = (hf - hf.mean(-1, keepdim=True)) * torch.rsqrt(hf.var(-1, unbiased=False, keepdim=True) + 1e-5)
            ref_g += (dy.float() * xhat).sum(0); ref_b += dy.float().sum(0)
            d2 = torch.randn(rows, cols, device="cuda").to(torch.bfloat16)
            assert cuda_ops.bias_grad(bias, d2) is None
            ref_bias += d2.float().sum(0)
    finally:
        ops.set_fused_wgrad_accumulation(False)
    assert _rel(g.main_grad, ref_g) < 2e-3 and _rel(b.main_grad, ref_b) < 2e-3 and _rel(bias.main_grad, ref_bias) < 1e-4
