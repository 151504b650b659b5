"""LayerNorm forward / backward kernel times (rows 16384 x cols 1024 bf16), fused vs split backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quintnet_b200 import ops
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
from torch.profiler import profile, ProfilerActivity
for rows, cols in ((16384, 1024), (8192, 1024), (16384, 768)):
    x = torch.randn(rows, cols, device=dev, dtype=torch.bfloat16, requires_grad=True)
    r = torch.randn(rows, cols, device=dev, dtype=torch.bfloat16, requires_grad=True)
    w = torch.randn(cols, device=dev, dtype=torch.bfloat16, requires_grad=True)
    b = torch.randn(cols, device=dev, dtype=torch.bfloat16, requires_grad=True)
    for mode in ("fused", "split"):
        if mode == "split": os.environ["QN_LN_BWD_SPLIT"] = "1"
        else: os.environ.pop("QN_LN_BWD_SPLIT", None)
        def step():
            flush.zero_()
            y, h = ops.dropout_add_layer_norm(x, r, w, b, 1e-5, 0.1, True)
            torch.autograd.backward([y, h], [y.detach(), h.detach()])
        for _ in range(3): step()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10): step()
            torch.cuda.synchronize()
        mb = rows * cols * 2 / 1e6
        for e in prof.key_averages():
            if "qn::" in e.key and e.device_time_total > 0:
                print(f"{rows}x{cols} {mode:5s} {e.key[:70]:70s} {e.device_time_total / e.count:7.1f} us x{e.count}  (one tensor = {mb:.0f} MB = {mb / 6.5729:.1f} us at HBM peak)")
