"""Stand-alone timing of the fused TP kernels (AG->GEMM, GEMM->RS) at tp=2 for the GPT-2 medium layer shapes,
with the QN_TP_DEBUG attribution knobs.  torchrun --nproc-per-node 2 scripts/bench_tp_kernels.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from quintnet_b200.ops import tp_fused, cuda_ops
from quintnet_b200.ops.cuda_ops import EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_MUL_AUX

g = dist.group.WORLD
ws = tp_fused._ws(g, dev)
M = 16384; rows = M // world
E = 1024
def rnd(*s): return (torch.randn(*s, device=dev) * 0.05).to(torch.bfloat16)

# (name, kind, N, K, b_mn, epilogue)
CASES = [
    ("ag c_attn fwd      N1536 K1024", "ag", 3 * E // world, E, False, EPI_BIAS),
    ("ag fc fwd +gelu    N2048 K1024", "ag", 4 * E // world, E, False, EPI_BIAS_GELU),
    ("ag c_proj dgrad    N512  K1024", "ag", E // world, E, True, EPI_NONE),
    ("ag mlpproj dgrad*g N2048 K1024", "ag", 4 * E // world, E, True, EPI_MUL_AUX),
    ("rs c_proj fwd      N1024 K512 ", "rs", E, E // world, False, EPI_BIAS),
    ("rs mlpproj fwd     N1024 K2048", "rs", E, 4 * E // world, False, EPI_BIAS),
    ("rs c_attn dgrad    N1024 K1536", "rs", E, 3 * E // world, True, EPI_NONE),
    ("rs fc dgrad        N1024 K2048", "rs", E, 4 * E // world, True, EPI_NONE),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run_case(kind, N, K, b_mn, epi, iters=12):
    if kind == "ag":
        x = rnd(rows, K)
        w = rnd(K, N) if b_mn else rnd(N, K)
        bias = rnd(N) if epi in (EPI_BIAS, EPI_BIAS_GELU) else None
        z = rnd(M, N) if epi in (EPI_BIAS_GELU, EPI_MUL_AUX) else None
        fn = lambda: tp_fused._ag_gemm(ws, x, w, b_mn, N, K, epi, bias, z, True, transient=b_mn)
    else:
        a = rnd(M, K)
        w = rnd(K, N) if b_mn else rnd(N, K)
        bias = rnd(N) if epi == EPI_BIAS else None
        fn = lambda: tp_fused._gemm_rs(ws, a, w, b_mn, N, K, bias)
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(iters):
        flush.zero_()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    t = torch.tensor([ts[len(ts) // 2]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # back-to-back (as in the model: no host sync between launches, ranks in lock-step through the flags)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    tb = torch.tensor([e0.elapsed_time(e1) * 1e3 / iters], device=dev)
    dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    return t.item(), tb.item()

out = {}
# unfused baselines: the same GEMM shapes through the plain kernel (no communication)
for name, kind, N, K, b_mn, epi in CASES:
    a = rnd(M, K); w = rnd(K, N) if b_mn else rnd(N, K)
    for _ in range(3): cuda_ops.gemm(a, w, M, N, K, b_mn=b_mn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(12): cuda_ops.gemm(a, w, M, N, K, b_mn=b_mn)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{name} plain GEMM (no comm, no epilogue) back-to-back {e0.elapsed_time(e1) * 1e3 / 12:6.1f} us  dbg=-", flush=True)
# mode:pairs[:dbg]  — dbg = QN_TP_DEBUG attribution knobs (results WRONG when set): AG 4 skip waits, 8 skip push, 12 both;
# RS 1 skip the peer wait + add
MODES = [m for m in os.environ.get("QN_BENCH_TP_MODES", "pump:1,pump:1:4,pump:1:8,pump:1:12,pump:1:1").split(",") if m]
for spec in MODES:
    parts = spec.split(":")
    mode, pairs = parts[0], parts[1]
    os.environ["QN_TP_AG_MODE"] = mode
    os.environ["QN_TP_2CTA"] = pairs
    for dbg in (int(parts[2]) if len(parts) > 2 else 0,):
        os.environ["QN_TP_DEBUG"] = str(dbg)
        for name, kind, N, K, b_mn, epi in CASES:
            if dbg and ((kind == "ag") != bool(dbg & 12)):
                continue                                   # knob does not apply to this kernel family
            t, tb = run_case(kind, N, K, b_mn, epi)
            out[f"{name} ag={mode} pairs={pairs} dbg={dbg}"] = {"us_isolated": round(t, 1), "us_back_to_back": round(tb, 1)}
            if rank == 0:
                print(f"{name} ag={mode:4s} pairs={pairs} dbg={dbg:2d}: isolated {t:6.1f} us  back-to-back {tb:6.1f} us", flush=True)
os.environ["QN_TP_AG_MODE"] = "pump"
os.environ["QN_TP_2CTA"] = "1"
os.environ["QN_TP_DEBUG"] = "0"
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(os.environ.get("QN_BENCH_TP_OUT", "gpurun_out/r2_tp_kernel_times.json"), "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
