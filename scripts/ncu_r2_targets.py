"""Launch every hot kernel ONCE inside a cudaProfilerStart/Stop window (after warm-up) at the benchmark shapes, for
`ncu --profile-from-start off --set full`.  The fused tensor-parallel kernels run as rank 0 of a loop-back world of 2
with QN_TP_DEBUG waits disabled, so a single launch completes under ncu's serialised kernel replay (the copy pump /
staged-slab stores execute for real against local "peer" buffers; results are not checked here).
usage: python scripts/ncu_r2_targets.py [--time]"""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quintnet_b200.ops import cuda_ops, native, optim_ops, tp_fused  # noqa: F401

dev = torch.device("cuda")
torch.manual_seed(0)
lib = native.require()
M, E = 16384, 1024
bf = torch.bfloat16
rnd = lambda *s: torch.randn(*s, device=dev, dtype=bf)
targets = []

# --- GEMMs (CTA pairs are the default for these shapes)
a, w, b = rnd(M, E), rnd(4 * E, E), rnd(4 * E)
z = torch.empty(M, 4 * E, device=dev, dtype=bf)
dy = rnd(M, 4 * E)
targets.append(("gemm_fwd_16384x4096x1024", lambda: cuda_ops.gemm(a, w, M, 4 * E, E), 2.0 * M * 4 * E * E, 0))
targets.append(("gemm_bias_gelu_16384x4096x1024", lambda: cuda_ops.gemm(a, w, M, 4 * E, E, epilogue=cuda_ops.EPI_BIAS_GELU, bias=b, z=z), 2.0 * M * 4 * E * E, 0))
targets.append(("gemm_dgrad_16384x1024x4096", lambda: cuda_ops.gemm(dy, w, M, E, 4 * E, b_mn=True), 2.0 * M * 4 * E * E, 0))
mg = torch.zeros(4 * E, E, device=dev)
wp = torch.nn.Parameter(w.clone()); wp.main_grad = mg
def _wg():
    from quintnet_b200 import ops
    ops.set_fused_wgrad_accumulation(True)
    cuda_ops.wgrad(wp, dy, a, 4 * E, E, M)
    ops.set_fused_wgrad_accumulation(False)
targets.append(("gemm_wgrad_accum_f32_4096x1024x16384", _wg, 2.0 * M * 4 * E * E, 0))

# --- attention
B, H, T = 32, 16, 512
qkv = rnd(B, T, 3 * H * 64).requires_grad_(True)
out = cuda_ops.attention_qkvpacked(qkv, H, True, 0.1)
dout = torch.randn_like(out)
fl_f = 4.0 * B * H * T * T * 64 / 2
targets.append(("attn_fwd_B32H16T512", lambda: cuda_ops.attention_qkvpacked(qkv.detach(), H, True, 0.1), fl_f, 0))
targets.append(("attn_bwd_B32H16T512", lambda: out.backward(dout, retain_graph=True), fl_f * 2.5, 0))

# --- LayerNorm family (bytes = algorithmic HBM traffic)
x, resid = rnd(M, E), rnd(M, E)
g, be = rnd(E).requires_grad_(True), rnd(E).requires_grad_(True)
xr = x.clone().requires_grad_(True); rr = resid.clone().requires_grad_(True)
yln = cuda_ops.dropout_add_layer_norm(xr, rr, g, be, 1e-5, 0.1)
ys = yln if isinstance(yln, (tuple, list)) else (yln,)
dys = [torch.randn_like(t) for t in ys]
targets.append(("dropout_add_layernorm_fwd_16384x1024", lambda: cuda_ops.dropout_add_layer_norm(x, resid, g.detach(), be.detach(), 1e-5, 0.1), 0, 4 * M * E * 2))
targets.append(("dropout_add_layernorm_bwd_16384x1024", lambda: torch.autograd.backward(ys, dys, retain_graph=True), 0, 5 * M * E * 2))

# --- AdamW (bf16 params, fp32 main grads, fp32 master/m/v) on a 64 Mi-element bucket
n = 64 << 20
p_, g_ = torch.zeros(n, device=dev, dtype=bf), torch.randn(n, device=dev)
m_, v_, ma_ = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
targets.append(("adamw_64Mi_bf16param_f32grad", lambda: optim_ops.adamw_flat(p_, g_, m_, v_, ma_, step=1, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01), 0, n * (4 + 12 + 12 + 2)))

# --- cross-entropy on bf16 logits
V = 50257
logits = rnd(8192, 50304)
labels = torch.randint(0, V, (8192,), device=dev)
targets.append(("ce_fwd_bwd_8192x50257", lambda: cuda_ops._ce_inplace(logits, labels, V, -100, True), 0, 2 * 8192 * V * 2))

# --- fused TP kernels, rank 0 of a loop-back world of 2 (tp2 shapes of GPT-2 medium)
ptrs = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
rows = M // 2
xl = rnd(rows, E)
gather = [torch.zeros(M * E, device=dev, dtype=bf) for _ in range(2)]
agf = [torch.zeros(1 << 18, device=dev, dtype=torch.int32) for _ in range(2)]
w_ag = rnd(1536, E); b_ag = rnd(1536); y_ag = torch.empty(M, 1536, device=dev, dtype=bf)
seq = [0]
def _ag(mode):
    os.environ["QN_TP_DEBUG"] = "4"; os.environ["QN_TP_AG_MODE"] = mode
    seq[0] += 1
    rc = lib.qn_gemm_ag_bf16(xl.data_ptr(), ptrs(gather), ptrs(agf), 0, 2, seq[0], w_ag.data_ptr(), y_ag.data_ptr(), M, 1536, E,
                             E, 1536, 0, 1, b_ag.data_ptr(), None, 0, 0, 0, None, native.stream())
    assert rc == 0, rc
    os.environ["QN_TP_DEBUG"] = "0"
targets.append(("fused_ag_gemm_pump_16384x1536x1024", lambda: _ag("pump"), 2.0 * M * 1536 * E, 0))
targets.append(("fused_ag_gemm_lsu_16384x1536x1024", lambda: _ag("lsu"), 2.0 * M * 1536 * E, 0))
a_rs = rnd(M, 2048); w_rs = rnd(E, 2048); y_rs = torch.empty(rows, E, device=dev, dtype=bf)
stage = [torch.zeros(M * E * 2, device=dev, dtype=torch.uint8) for _ in range(2)]
rsf = [torch.zeros(1 << 18, device=dev, dtype=torch.int32) for _ in range(2)]
def _rs():
    os.environ["QN_TP_DEBUG"] = "1"
    seq[0] += 1
    rc = lib.qn_gemm_rs_bf16(a_rs.data_ptr(), w_rs.data_ptr(), y_rs.data_ptr(), ptrs(stage), ptrs(rsf), 0, 2, seq[0], M, E, 2048,
                             2048, 2048, E, 0, None, 0, 0, native.stream())
    assert rc == 0, rc
    os.environ["QN_TP_DEBUG"] = "0"
targets.append(("fused_gemm_rs_16384x1024x2048", _rs, 2.0 * M * E * 2048, 0))
targets.append(("plain_gemm_16384x1536x1024", lambda: cuda_ops.gemm(rnd(M, E), w_ag, M, 1536, E), 2.0 * M * 1536 * E, 0))
targets.append(("plain_gemm_16384x1024x2048", lambda: cuda_ops.gemm(a_rs, w_rs, M, E, 2048), 2.0 * M * E * 2048, 0))

for _, fn, _, _ in targets:          # warm-up (module load, attribute set-up, allocator)
    fn(); fn()
torch.cuda.synchronize()
if "--time" in sys.argv:
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}
    for name, fn, flops, nbytes in targets:
        ts = []
        for _ in range(15):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort(); us = ts[len(ts) // 2]
        res[name] = {"us": round(us, 1), "tflops": round(flops / us / 1e6, 1) if flops else None,
                     "gbs": round(nbytes / us / 1e3, 1) if nbytes else None}
        print(f"{name:45s} {us:9.1f} us  {res[name]['tflops'] or ''} TF  {res[name]['gbs'] or ''} GB/s", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/r2_kernel_times.json", "w"), indent=1)
else:
    torch.cuda.profiler.start()
    for name, fn, _, _ in targets:
        torch.cuda.nvtx.range_push(name); fn(); torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
