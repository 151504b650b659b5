"""Tiny driver that launches each hot kernel a few times at the benchmark shapes (for ncu captures and
stand-alone CUDA-event timing).  usage: python scripts/ncu_targets.py [gemm|attn|all] [--time]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quintnet_b200.ops import cuda_ops

what = sys.argv[1] if len(sys.argv) > 1 else "all"
do_time = "--time" in sys.argv
dev = torch.device("cuda")
torch.manual_seed(0)
res = {}


def timeit(name, fn, flops, iters=20):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                      # evict L2 between timed iterations
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    med = ts[len(ts) // 2]
    res[name] = {"ms": med, "tflops": flops / med / 1e9}
    print(f"{name:40s} {med*1e3:9.1f} us   {flops/med/1e9:8.1f} TFLOP/s", flush=True)


if what in ("gemm", "all"):
    M = 16384
    for (N, K, tag) in [(4096, 1024, "c_fc fwd"), (1024, 4096, "mlp c_proj fwd"), (3072, 1024, "c_attn fwd"), (1024, 1024, "attn c_proj fwd"), (1536, 1024, "tp2 c_attn"), (1024, 512, "tp2 attn c_proj")]:
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        dy = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        if do_time:
            timeit(f"gemm TN {M}x{N}x{K} ({tag})", lambda: cuda_ops.gemm(a, w, M, N, K), fl)
            timeit(f"gemm dgrad {M}x{K}x{N}", lambda: cuda_ops.gemm(dy, w, M, K, N, b_mn=True), fl)
            timeit(f"gemm wgrad {N}x{K}x{M}", lambda: cuda_ops.gemm(dy, a, N, K, M, a_mn=True, b_mn=True), fl)
            timeit(f"gemm wgrad {N}x{K}x{M} no-splitK", lambda: cuda_ops.gemm(dy, a, N, K, M, a_mn=True, b_mn=True, split_k=1), fl)
            timeit(f"cuBLAS TN {M}x{N}x{K}", lambda: torch.matmul(a, w.t()), fl)
        else:
            cuda_ops.gemm(a, w, M, N, K); cuda_ops.gemm(dy, w, M, K, N, b_mn=True); cuda_ops.gemm(dy, a, N, K, M, a_mn=True, b_mn=True)
    if do_time:
        z = torch.randn(M, 4096, device=dev, dtype=torch.bfloat16)
        a = torch.randn(M, 1024, device=dev, dtype=torch.bfloat16); w = torch.randn(4096, 1024, device=dev, dtype=torch.bfloat16)
        b = torch.randn(4096, device=dev, dtype=torch.bfloat16)
        zz = torch.empty(M, 4096, device=dev, dtype=torch.bfloat16)
        timeit("gemm c_fc bias+gelu (+z)", lambda: cuda_ops.gemm(a, w, M, 4096, 1024, epilogue=cuda_ops.EPI_BIAS_GELU, bias=b, z=zz), 2.0 * M * 4096 * 1024)
        dy = torch.randn(M, 1024, device=dev, dtype=torch.bfloat16); w2 = torch.randn(1024, 4096, device=dev, dtype=torch.bfloat16)
        timeit("gemm dgrad * aux epilogue", lambda: cuda_ops.gemm(dy, w2, M, 4096, 1024, b_mn=True, epilogue=cuda_ops.EPI_MUL_AUX, z=z), 2.0 * M * 4096 * 1024)

if what in ("attn", "all"):
    B, H, T = 32, 16, 512
    qkv = torch.randn(B, T, 3 * H * 64, device=dev, dtype=torch.bfloat16, requires_grad=True)
    out = cuda_ops.attention_qkvpacked(qkv, H, True, 0.1)
    dout = torch.randn_like(out)
    out.backward(dout)
    fl_f = 4.0 * B * H * T * T * 64 / 2
    if do_time:
        for pd in (0.0, 0.1):
            timeit(f"attn fwd causal p={pd}", lambda: cuda_ops.attention_qkvpacked(qkv.detach(), H, True, pd), fl_f)
            o = cuda_ops.attention_qkvpacked(qkv, H, True, pd)
            timeit(f"attn bwd causal p={pd}", lambda: o.backward(dout, retain_graph=True), fl_f * 2.5)
        import torch.nn.functional as F
        q, k, v = qkv.detach().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        timeit("cuDNN/SDPA fwd causal p=0.1", lambda: F.scaled_dot_product_attention(q, k, v, is_causal=True, dropout_p=0.1), fl_f)
torch.cuda.synchronize()
if do_time:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/kernel_times_{what}.json", "w"), indent=1)
