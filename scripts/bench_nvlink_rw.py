"""NVLink in-kernel PUSH (remote stores) vs PULL (remote loads) bandwidth at several transfer sizes, so that the fixed
launch/sync overhead separates from the link rate (torchrun --nproc-per-node 2).  Same kernels for both directions:
`qn_push_bench` (LDG.128 -> STG.128, 16 in flight per thread) and `qn_push_bench_bulk` (1-D bulk async copies through
shared memory) — pull just swaps which side is the peer mapping."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from quintnet_b200.ops import symm, native
MAXB = 256 << 20
buf = symm.alloc_bytes(2 * MAXB, dev, dist.group.WORLD)          # [0, MAXB): data others read / write; [MAXB, 2 MAXB): local dst
peer = buf.peer_ptrs[(rank + 1) % world]
local = buf.peer_ptrs[rank]
lib = native.require()
buf.tensor.random_(0, 255)

def timed(fn, iters=6):
    ts = []
    for it in range(iters + 2):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); t = torch.tensor([ts[len(ts) // 2]], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

rows = []
for nbytes in (1 << 20, 4 << 20, 16 << 20, 64 << 20, 256 << 20):
    def st(src, dst):
        return lambda: native.check(lib.qn_push_bench(src, dst, nbytes, 296, 256, 16, 0, 0, native.stream()), "push")
    def bulk(src, dst, blocks=148, chunk=8192):
        return lambda: native.check(lib.qn_push_bench_bulk(src, dst, nbytes, blocks, chunk, native.stream()), "bulk")
    r = {"bytes": nbytes,
         "lsu_push_us": timed(st(local, peer + MAXB)), "lsu_pull_us": timed(st(peer, local + MAXB)),
         "lsu_local_us": timed(st(local, local + MAXB)),
         "bulk_push_us": timed(bulk(local, peer + MAXB)), "bulk_pull_us": timed(bulk(peer, local + MAXB)),
         "bulk_local_us": timed(bulk(local, local + MAXB))}
    rows.append(r)
    if rank == 0:
        print(f"{nbytes >> 20:4d} MiB  LSU push {r['lsu_push_us']:8.1f} us  pull {r['lsu_pull_us']:8.1f}  local {r['lsu_local_us']:8.1f} | "
              f"bulk push {r['bulk_push_us']:8.1f}  pull {r['bulk_pull_us']:8.1f}  local {r['bulk_local_us']:8.1f}", flush=True)
if rank == 0:
    # link rate from the slope between the two largest sizes (fixed overhead cancels)
    a, b = rows[-2], rows[-1]
    for k in ("lsu_push_us", "lsu_pull_us", "bulk_push_us", "bulk_pull_us"):
        gbs = (b["bytes"] - a["bytes"]) / (b[k] - a[k]) / 1e3
        print(f"{k[:-3]:10s} slope bandwidth {gbs:7.1f} GB/s per direction (both ranks active)", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/r2_nvlink_push_pull.json", "w"), indent=1)
dist.barrier(); dist.destroy_process_group()
