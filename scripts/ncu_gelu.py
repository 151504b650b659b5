import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quintnet_b200.ops import cuda_ops
M = 16384
a = torch.randn(M, 1024, device="cuda", dtype=torch.bfloat16); w = torch.randn(4096, 1024, device="cuda", dtype=torch.bfloat16)
b = torch.randn(4096, device="cuda", dtype=torch.bfloat16); z = torch.empty(M, 4096, device="cuda", dtype=torch.bfloat16)
for _ in range(45):
    cuda_ops.gemm(a, w, M, 4096, 1024, epilogue=cuda_ops.EPI_BIAS_GELU, bias=b, z=z)
torch.cuda.synchronize()
