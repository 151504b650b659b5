"""NVLink push bandwidth from inside a kernel vs launch shape, and its interference with a concurrent GEMM
(torchrun --nproc-per-node 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from quintnet_b200.ops import symm, native, cuda_ops
nbytes = 16 << 20
buf = symm.alloc_bytes(4 * nbytes, dev, dist.group.WORLD)
src = torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=dev)
peer = buf.peer_ptrs[(rank + 1) % world]
local = buf.peer_ptrs[rank]
lib = native.require()

def timed(fn, both=True, iters=8):
    ts = []
    for it in range(iters + 2):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if both or rank == 0:
            fn()
        e1.record(); torch.cuda.synchronize()
        if it >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); t = torch.tensor([ts[len(ts) // 2]], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

def st(dst, blocks, threads, depth, seg=0, stride=0):
    def f():
        rc = lib.qn_push_bench(src.data_ptr(), dst, nbytes, blocks, threads, depth, seg, stride, native.stream()); assert rc == 0, rc
    return f
def bulk(dst, blocks, chunk):
    def f():
        rc = lib.qn_push_bench_bulk(src.data_ptr(), dst, nbytes, blocks, chunk, native.stream()); assert rc == 0, rc
    return f
def show(label, t):
    if rank == 0: print(f"{label}: {t:7.1f} us  {nbytes / t / 1e3:7.1f} GB/s", flush=True)

if rank == 0: print("16 MiB push; both ranks push simultaneously unless noted; GB/s per direction")
for blocks, threads, depth in ((8, 512, 8), (16, 512, 8), (32, 512, 8), (148, 256, 8), (296, 256, 16)):
    show(f"st.v4  peer  {blocks:3d} x {threads:3d} thr depth {depth:2d}", timed(st(peer, blocks, threads, depth)))
show("st.v4  peer  296 x 256 one-directional    ", timed(st(peer, 296, 256, 16), both=False))
show("st.v4  local 296 x 256                    ", timed(st(local, 296, 256, 16)))
for blocks in (4, 8, 16, 32, 148):
    for chunk in (8192, 32768):
        show(f"bulk   peer  {blocks:3d} CTAs chunk {chunk:5d}        ", timed(bulk(peer, blocks, chunk)))
show("bulk   peer   32 CTAs chunk 32768 one-dir ", timed(bulk(peer, 32, 32768), both=False))
show("bulk   local  32 CTAs chunk 32768         ", timed(bulk(local, 32, 32768)))

# interference: plain GEMM (16384 x 1536 x 1024) on 148-P CTAs while P CTAs push on a second stream
M, N, K = 16384, 1536, 1024
a = (torch.randn(M, K, device=dev) * 0.05).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
side = torch.cuda.Stream()
def gemm_only(ctas):
    return lambda: cuda_ops.gemm(a, w, M, N, K, max_ctas=ctas)
def gemm_with(ctas, pushfn):
    def f():
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            pushfn()
        cuda_ops.gemm(a, w, M, N, K, max_ctas=ctas)
        torch.cuda.current_stream().wait_stream(side)
    return f
def showg(label, t):
    if rank == 0: print(f"{label}: {t:7.1f} us", flush=True)
showg("gemm 148 CTAs alone                        ", timed(gemm_only(148)))
showg("gemm 132 CTAs alone                        ", timed(gemm_only(132)))
showg("gemm 140 CTAs alone                        ", timed(gemm_only(140)))
showg("gemm 132 CTAs + st.v4 push on 16 x 512     ", timed(gemm_with(132, st(peer, 16, 512, 8))))
showg("gemm 132 CTAs + bulk push on 16 CTAs 32K   ", timed(gemm_with(132, bulk(peer, 16, 32768))))
showg("gemm 140 CTAs + bulk push on 8 CTAs 32K    ", timed(gemm_with(140, bulk(peer, 8, 32768))))
showg("gemm 148 CTAs + bulk push on 8 CTAs 32K    ", timed(gemm_with(148, bulk(peer, 8, 32768))))
showg("gemm 148 CTAs + st.v4 push on 148 x 64     ", timed(gemm_with(148, st(peer, 148, 64, 16))))
dist.barrier(); dist.destroy_process_group()
