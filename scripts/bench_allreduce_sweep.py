"""BASELINE config #5: gradient-bucket all-reduce bandwidth sweep 1 KB - 1 GB, K1 (in-kernel NVLS / P2P over symmetric
memory; one-shot for small messages, two-shot for large) vs ncclAllReduce, on all ranks of the launch.
  torchrun --nproc-per-node N scripts/bench_allreduce_sweep.py
Reports per size: time (device events, max over ranks), algbw = bytes/time, busbw = algbw * 2(N-1)/N.
Reference call site replaced: parallelism/data_parallel/components/gradient_reducer.py:87-96 -> backends/torch_backend.py:71."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from quintnet_b200.ops import symm

g = dist.group.WORLD
MAXB = int(os.environ.get("QN_SWEEP_MAX_BYTES", str(1 << 30)))
sizes = [1 << k for k in range(10, 31) if (1 << k) <= MAXB]
buf = symm.alloc(MAXB // 2, torch.bfloat16, dev, g)                      # symmetric bf16 gradient bucket
nccl_buf = torch.empty(MAXB // 2, dtype=torch.bfloat16, device=dev)
b_off = symm.lookup(buf)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters, flush_l2):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(iters):
        if flush_l2:
            flush.zero_()
        dist.barrier(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    t = torch.tensor([ts[len(ts) // 2]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def timed_b2b(fn, iters):
    """back to back (how a bucketed backward issues them): launch overhead amortised"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) * 1e3 / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


rows = []
for nbytes in sizes:
    n = nbytes // 2
    view = buf[:n]
    nv = nccl_buf[:n]
    iters = 20 if nbytes <= (64 << 20) else 6
    res = {"bytes": nbytes}
    res["nccl_us"] = timed_b2b(lambda: dist.all_reduce(nv), iters)
    if n % (world * 8) == 0:
        res["k1_nvls_us"] = timed_b2b(lambda: symm.all_reduce_(b_off, view, 1.0, use_multicast=True), iters) if b_off[0].mc_ptr else None
        res["k1_p2p_us"] = timed_b2b(lambda: symm.all_reduce_(b_off, view, 1.0, use_multicast=False), iters)
    if nbytes <= symm.ONESHOT_MAX_BYTES:
        x = nv
        res["k1_oneshot_us"] = timed_b2b(lambda: symm.all_reduce_small(x, g), iters)
    best = min(v for k, v in res.items() if k.startswith("k1_") and v)
    res["k1_best_us"] = best
    f = 2.0 * (world - 1) / world
    res["nccl_busbw_gbs"] = nbytes / res["nccl_us"] / 1e3 * f
    res["k1_busbw_gbs"] = nbytes / best / 1e3 * f
    rows.append(res)
    if rank == 0:
        print(f"{nbytes:>11d} B  nccl {res['nccl_us']:9.1f} us ({res['nccl_busbw_gbs']:7.1f} GB/s)  "
              f"k1 nvls {res.get('k1_nvls_us') or 0:9.1f}  p2p {res.get('k1_p2p_us') or 0:9.1f}  one-shot {res.get('k1_oneshot_us') or 0:8.1f}  "
              f"-> best {best:9.1f} us ({res['k1_busbw_gbs']:7.1f} GB/s busbw)  k1/nccl speed {res['nccl_us'] / best:5.2f}x", flush=True)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"world": world, "dtype": "bf16", "timing": "back-to-back launches, CUDA events, max over ranks",
               "multicast": bool(b_off[0].mc_ptr), "rows": rows}, open(f"gpurun_out/r2_allreduce_sweep_{world}gpu.json", "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
