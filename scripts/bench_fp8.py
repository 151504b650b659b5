"""FP8 (e4m3, row-wise scales) GEMM vs the bf16 tcgen05 GEMM on the GPT-2 medium forward shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quintnet_b200.ops import cuda_ops
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=20):
    ts = []
    for _ in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[3:]); return ts[len(ts) // 2]
for M, N, K in ((16384, 4096, 1024), (16384, 1024, 4096), (16384, 3072, 1024), (8192, 8192, 8192)):
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    qa, sa = cuda_ops.quantize_rowwise(a); qw, sw = cuda_ops.quantize_rowwise(w)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t8 = timeit(lambda: cuda_ops.gemm_fp8(qa, sa, qw, sw, M, N, K, out=out))
    t16 = timeit(lambda: cuda_ops.gemm(a, w, M, N, K, out=out))
    tq = timeit(lambda: cuda_ops.quantize_rowwise(a))
    ma, sfa = cuda_ops.quantize_mx(a); mw, sfw = cuda_ops.quantize_mx(w)
    tmx = timeit(lambda: cuda_ops.gemm_mxfp8(ma, sfa, mw, sfw, M, N, K))
    tqmx = timeit(lambda: cuda_ops.quantize_mx(a))
    fl = 2.0 * M * N * K
    print(f"{M}x{N}x{K}: mxfp8 block-scaled {tmx:6.1f} us ({fl / tmx / 1e6:6.0f} TFLOP/s)  fp8 row-wise {t8:6.1f} us ({fl / t8 / 1e6:6.0f} TFLOP/s)  "
          f"bf16 {t16:6.1f} us ({fl / t16 / 1e6:6.0f} TFLOP/s)  quantize A: mx {tqmx:5.1f} us / row-wise {tq:5.1f} us", flush=True)
