"""Round-2 experiment: does NVLink traffic by itself slow a concurrent tensor-core GEMM?

Each rank runs the c_attn-shaped GEMM (16384 x 1536 x 1024, CTA pairs) back to back on the compute stream, alone and
while a side stream keeps pushing 16 MiB to the peer's symmetric buffer with (a) cudaMemcpyAsync (copy engine, no SM
work) and (b) the SM push kernel on 16 CTAs.  If (a) leaves the GEMM time unchanged, the fused AG kernels should
hand their all-gather to the copy engine (or to the producing LayerNorm) instead of push warps.

    torchrun --nproc-per-node 2 scripts/exp_copy_engine_overlap.py
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from quintnet_b200.ops import cuda_ops, native, symm
cudart = ctypes.CDLL("libcudart.so")
nbytes = 16 << 20
buf = symm.alloc_bytes(nbytes, dev, dist.group.WORLD)
src = torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=dev)
peer = buf.peer_ptrs[(rank + 1) % world]
lib = native.require()
M, N, K = 16384, 1536, 1024
a = (torch.randn(M, K, device=dev) * 0.05).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
side = torch.cuda.Stream()
REPS = 20

def gemms():
    for _ in range(REPS):
        cuda_ops.gemm(a, w, M, N, K, out=out)

def measure(push):
    """GEMM time per launch on the compute stream while `push` (or nothing) loops on the side stream."""
    for _ in range(2):
        gemms()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if push is not None:
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(REPS * 3):
                push(side.cuda_stream)
    e0.record(); gemms(); e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / REPS], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

def push_ce(stream):
    rc = cudart.cudaMemcpyAsync(ctypes.c_void_p(peer), ctypes.c_void_p(src.data_ptr()), ctypes.c_size_t(nbytes), 3,
                                ctypes.c_void_p(stream))
    assert rc == 0, rc

def push_sm(stream):
    rc = lib.qn_push_bench(src.data_ptr(), peer, nbytes, 16, 512, 8, 0, 0, ctypes.c_void_p(stream))
    assert rc == 0, rc

for name, fn in (("alone", None), ("+ copy-engine NVLink push", push_ce), ("+ SM push kernel on 16 CTAs", push_sm)):
    t = measure(fn)
    if rank == 0:
        print(f"GEMM {M}x{N}x{K} {name:32s}: {t:6.1f} us per launch", flush=True)
dist.barrier(); dist.destroy_process_group()
