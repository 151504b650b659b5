"""Sustained (power-capped) GEMM throughput, ours vs cuBLAS, apples to apples: each implementation runs back to back
for ~2 s on fresh-from-L2-evicted operands of the benchmark shapes; clocks / power sampled with nvidia-smi meanwhile.
MEASURED_PEAKS.json's `bf16_tflops_sustained` is the same protocol on 8192^3."""
import json, os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quintnet_b200.ops import cuda_ops

dev = torch.device("cuda")
torch.manual_seed(0)
SHAPES = [(16384, 4096, 1024), (16384, 1024, 4096), (16384, 3072, 1024), (8192, 8192, 8192)]
res = {}


def smi():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True).stdout.strip().split(",")
    return float(out[0]), float(out[1])


def sustained(fn, flops, seconds=2.0):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    samples = []
    stop = threading.Event()
    def sampler():
        while not stop.is_set():
            try:
                samples.append(smi())
            except Exception:
                pass
            time.sleep(0.1)
    th = threading.Thread(target=sampler); th.start()
    n = 0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    s.record()
    while time.time() - t0 < seconds:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    ms = s.elapsed_time(e)
    clk = sorted(x[0] for x in samples)[len(samples) // 2] if samples else None
    pw = max(x[1] for x in samples) if samples else None
    return {"us": ms * 1e3 / n, "tflops": flops * n / ms / 1e9, "sm_mhz_median": clk, "power_w_max": pw}


for (M, N, K) in SHAPES:
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    r_ours = sustained(lambda: cuda_ops.gemm(a, w, M, N, K, out=out), fl)
    r_lib = sustained(lambda: torch.matmul(a, w.t(), out=out), fl)
    res[f"{M}x{N}x{K}"] = {"ours": r_ours, "cublas": r_lib}
    print(f"{M}x{N}x{K}: ours {r_ours['us']:.1f} us {r_ours['tflops']:.0f} TF @ {r_ours['sm_mhz_median']} MHz {r_ours['power_w_max']} W | "
          f"cuBLAS {r_lib['us']:.1f} us {r_lib['tflops']:.0f} TF @ {r_lib['sm_mhz_median']} MHz {r_lib['power_w_max']} W", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r2_gemm_sustained.json", "w"), indent=1)
