"""Plain tcgen05 GEMM: 1-CTA tiles vs CTA pairs (cta_group::2) vs cuBLAS on the GPT-2 medium shapes."""
import os, sys, subprocess
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
mode = os.environ.get("QN_GEMM_2CTA", "0")
for name, M, N, K, a_mn, b_mn in (("fwd  X.Wt", 16384, 4096, 1024, False, False), ("fwd  X.Wt", 16384, 1024, 4096, False, False),
                                   ("fwd  X.Wt", 16384, 3072, 1024, False, False), ("dgrad dY.W", 16384, 1024, 4096, False, True),
                                   ("wgrad dYt.X", 4096, 1024, 16384, True, True), ("square", 8192, 8192, 8192, False, False)):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev, dtype=torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device=dev, dtype=torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: cuda_ops.gemm(a, b, M, N, K, a_mn=a_mn, b_mn=b_mn, out=out))
    A = a.t() if a_mn else a; B = b if b_mn else b.t()
    ref = (A.float() @ B.float())
    err = ((out.float() - ref).norm() / ref.norm()).item()
    tc = timeit(lambda: torch.matmul(A, B, out=out))
    fl = 2.0 * M * N * K
    print(f"2cta={mode} {name:11s} {M}x{N}x{K}: ours {t:6.1f} us ({fl / t / 1e6:5.0f} TF)  cuBLAS {tc:6.1f} us ({fl / tc / 1e6:5.0f} TF)  rel err {err:.1e}")
