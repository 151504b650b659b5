"""In-kernel NVLink collectives (K1 all-reduce, K2 ZeRO-1 phases) vs NCCL, on >= 2 GPUs."""
import pytest
import torch
import torch.distributed as dist

from tests._dist import run_distributed

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu(2)]


def _allreduce(rank, world):
    from quintnet_b200.ops import symm
    dev = torch.device("cuda", rank)
    g = dist.group.WORLD
    for dtype in (torch.float32, torch.bfloat16):
        n = 1 << 20
        t = symm.alloc(n, dtype, dev, g)
        torch.manual_seed(rank)
        src = torch.randn(n, device=dev).to(dtype)
        for use_mc in (False, True):
            t.copy_(src)
            ref = src.float().clone()
            dist.all_reduce(ref)
            torch.cuda.synchronize()
            dist.barrier()
            symm.all_reduce_(symm.lookup(t), t, 1.0 / world, use_multicast=use_mc)
            torch.cuda.synchronize()
            tol = 1e-5 if dtype == torch.float32 else 2e-2
            err = ((t.float() - ref / world).abs().max() / ref.abs().max()).item()
            assert err < tol, (dtype, use_mc, err)
    if rank == 0:
        b, _ = symm.lookup(t)
        print(f"symm backend={b.backend} multicast={'yes' if b.mc_ptr else 'no'}")


def _oneshot(rank, world):
    """one-launch small all-reduce (NVLS / P2P one-shot, device-side launch counter) vs NCCL, many launches in a row"""
    from quintnet_b200.ops import symm
    dev = torch.device("cuda", rank)
    g = dist.group.WORLD
    for dtype, n in ((torch.bfloat16, 8 * 50 * 64), (torch.float32, 1), (torch.bfloat16, 100003), (torch.float32, 65536)):
        for it in range(5):
            torch.manual_seed(100 * rank + it)
            x = torch.randn(n, device=dev).to(dtype)
            ref = x.float().clone()
            dist.all_reduce(ref)
            y = symm.all_reduce_small(x, g, 1.0)
            tol = 1e-5 if dtype == torch.float32 else 2e-2
            err = ((y.float() - ref).abs().max() / ref.abs().max().clamp(min=1e-6)).item()
            assert err < tol, (dtype, n, it, err)
    # CUDA-graph replay: the launch counter lives on the device, so a captured launch stays in lock-step
    x = torch.randn(4096, device=dev).to(torch.bfloat16)
    symm.all_reduce_small(x, g)
    torch.cuda.synchronize(); dist.barrier()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        y = symm.all_reduce_small(x, g)
    for it in range(4):
        x.copy_(torch.randn(4096, device=dev).to(torch.bfloat16))
        ref = x.float().clone(); dist.all_reduce(ref)
        gr.replay()
        torch.cuda.synchronize()
        assert ((y.float() - ref).abs().max() / ref.abs().max()).item() < 2e-2


def _ddp_native(rank, world):
    """DataParallel with the NVLink backend == averaged reference gradients."""
    import torch.nn as nn
    from quintnet_b200 import DataParallel, init_process_groups
    from quintnet_b200.parallel.data_parallel import BucketConfig, DistributedConfig, NVLinkBackend
    pg = init_process_groups("cuda", [world], ["dp"])
    dev = torch.device("cuda", rank)
    torch.manual_seed(0)
    ref = nn.Sequential(nn.Linear(64, 256), nn.GELU(), nn.Linear(256, 64)).to(dev)
    torch.manual_seed(0)
    mod = nn.Sequential(nn.Linear(64, 256), nn.GELU(), nn.Linear(256, 64)).to(dev)
    ddp = DataParallel(mod, DistributedConfig(rank, world, pg.get_group("dp")), BucketConfig(capacity_mb=1),
                       backend=NVLinkBackend())
    torch.manual_seed(5)
    xs = torch.randn(world, 32, 64, device=dev)
    for r in range(world):
        (ref(xs[r]).pow(2).mean() / world).backward()
    ddp(xs[rank]).pow(2).mean().backward()
    torch.cuda.synchronize()
    for p, q in zip(ddp.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-5), (p.grad - q.grad).abs().max()


def _zero1_native(rank, world):
    """ZeRO-1 DistributedAdamW over symmetric buckets == torch AdamW on averaged grads."""
    import torch.nn as nn
    from quintnet_b200 import init_process_groups
    from quintnet_b200.optim import DistributedAdamW, wrap_zero1
    from quintnet_b200.parallel.data_parallel import NVLinkBackend
    pg = init_process_groups("cuda", [world], ["dp"])
    dev = torch.device("cuda", rank)
    for dtype in (torch.float32, torch.bfloat16):
        torch.manual_seed(0)
        ref = nn.Sequential(nn.Linear(64, 256), nn.GELU(), nn.Linear(256, 64)).to(dev)
        torch.manual_seed(0)
        mod = nn.Sequential(nn.Linear(64, 256), nn.GELU(), nn.Linear(256, 64)).to(dev).to(dtype)
        ddp = wrap_zero1(mod, pg.get_group("dp"), rank, world, backend=NVLinkBackend())
        opt = DistributedAdamW(ddp, lr=1e-2, weight_decay=0.01, max_grad_norm=1.0, pg_manager=pg)
        assert opt._use_native(opt.buckets[0]), "symmetric buckets expected"
        oref = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.01)
        torch.manual_seed(5)
        for step in range(3):
            xs = torch.randn(world, 32, 64, device=dev)
            oref.zero_grad()
            for r in range(world):
                (ref(xs[r]).pow(2).mean() / world).backward()
            torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
            oref.step()
            ddp(xs[rank].to(dtype)).float().pow(2).mean().backward()
            opt.step()
            opt.zero_grad()
        torch.cuda.synchronize()
        tol = 1e-4 if dtype == torch.float32 else 3e-2
        for p, q in zip(ddp.parameters(), ref.parameters()):
            assert torch.allclose(p.float(), q, atol=tol), (dtype, (p.float() - q).abs().max())


def test_allreduce_kernel():
    run_distributed(_allreduce, 2, backend="nccl")


def test_oneshot_small_allreduce():
    run_distributed(_oneshot, 2, backend="nccl")


def test_ddp_nvlink_backend():
    run_distributed(_ddp_native, 2, backend="nccl")


def test_zero1_fused_kernels():
    run_distributed(_zero1_native, 2, backend="nccl")


def _tp_fused(rank, world):
    """Fused AG->GEMM / GEMM->RS kernels == unfused collectives + dense GEMM, forward and backward."""
    import torch.nn.functional as F
    from quintnet_b200.ops import tp_fused
    dev = torch.device("cuda", rank)
    g = dist.group.WORLD
    torch.manual_seed(0)
    for (M, K, N) in [(512, 256, 384), (2048, 1024, 1536), (4096, 512, 256)]:
        rows = M // world
        X = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)             # same on every rank (same seed)
        W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
        b = torch.randn(N, device=dev).to(torch.bfloat16)
        dY = torch.randn(M, N, device=dev).to(torch.bfloat16)
        # ---- all-gather -> GEMM (+ gelu)
        for act in (None, "gelu"):
            xl = X[rank * rows:(rank + 1) * rows].clone().requires_grad_(True)
            w = W.clone().requires_grad_(True)
            bb = b.clone().requires_grad_(True)
            y = tp_fused.allgather_linear(xl, w, bb, g, act)
            y.backward(dY)
            Xr, Wr, br = X.float().requires_grad_(True), W.float().requires_grad_(True), b.float().requires_grad_(True)
            yr = F.linear(Xr, Wr, br)
            if act == "gelu":
                yr = F.gelu(yr)
            yr.backward(dY.float())
            rel = lambda a, c: ((a.float() - c.float()).norm() / (c.float().norm() + 1e-9)).item()
            assert rel(y, yr) < 2e-2, ("ag fwd", M, K, N, act, rel(y, yr))
            # dX_local = reduce-scatter over ranks of dY·W; every rank used the same dY -> world x the dense grad rows
            assert rel(xl.grad, Xr.grad[rank * rows:(rank + 1) * rows] * world) < 3e-2, ("ag dx", rel(xl.grad, Xr.grad[rank * rows:(rank + 1) * rows] * world))
            assert rel(w.grad, Wr.grad) < 3e-2, ("ag dw", rel(w.grad, Wr.grad))
            assert rel(bb.grad, br.grad) < 3e-2
        # ---- GEMM -> reduce-scatter (K sharded over ranks)
        kl = K // world
        xk = X[:, rank * kl:(rank + 1) * kl].contiguous().requires_grad_(True)
        wk = W[:, rank * kl:(rank + 1) * kl].contiguous().requires_grad_(True)
        bias = b.clone().requires_grad_(True) if rank == 0 else None
        yl = tp_fused.linear_reduce_scatter(xk, wk, bias, g)
        dYl = dY[rank * rows:(rank + 1) * rows]
        yl.backward(dYl)
        Xr, Wr = X.float().requires_grad_(True), W.float().requires_grad_(True)
        yr = F.linear(Xr, Wr, b.float())
        yr.backward(dY.float())
        assert rel(yl, yr[rank * rows:(rank + 1) * rows]) < 2e-2, ("rs fwd", M, K, N, rel(yl, yr[rank * rows:(rank + 1) * rows]))
        assert rel(xk.grad, Xr.grad[:, rank * kl:(rank + 1) * kl]) < 3e-2, ("rs dx",)
        assert rel(wk.grad, Wr.grad[:, rank * kl:(rank + 1) * kl]) < 3e-2, ("rs dw",)
    torch.cuda.synchronize()


def test_tp_fused_gemm_collectives():
    run_distributed(_tp_fused, 2, backend="nccl")


def _sp_mlp(rank, world):
    import torch.nn.functional as F
    from quintnet_b200.ops import tp_fused
    dev = torch.device("cuda", rank)
    g = dist.group.WORLD
    torch.manual_seed(0)
    M, E, I = 2048, 512, 2048
    rows, il = M // world, I // world
    X = (torch.randn(M, E, device=dev) * 0.5).to(torch.bfloat16)
    W1 = (torch.randn(I, E, device=dev) * 0.05).to(torch.bfloat16)
    b1 = (torch.randn(I, device=dev) * 0.1).to(torch.bfloat16)
    W2 = (torch.randn(E, I, device=dev) * 0.05).to(torch.bfloat16)
    b2 = (torch.randn(E, device=dev) * 0.1).to(torch.bfloat16)
    dY = torch.randn(M, E, device=dev).to(torch.bfloat16)
    xl = X[rank * rows:(rank + 1) * rows].clone().requires_grad_(True)
    w1 = W1[rank * il:(rank + 1) * il].clone().requires_grad_(True)
    bb1 = b1[rank * il:(rank + 1) * il].clone().requires_grad_(True)
    w2 = W2[:, rank * il:(rank + 1) * il].contiguous().requires_grad_(True)
    bb2 = b2.clone().requires_grad_(True) if rank == 0 else None
    y = tp_fused.sequence_parallel_mlp(xl, w1, bb1, w2, bb2, g)
    y.backward(dY[rank * rows:(rank + 1) * rows])
    Xr, W1r, b1r, W2r, b2r = (t.float().requires_grad_(True) for t in (X, W1, b1, W2, b2))
    yr = F.linear(F.gelu(F.linear(Xr, W1r, b1r)), W2r, b2r)
    yr.backward(dY.float())
    rel = lambda a, c: ((a.float() - c.float()).norm() / (c.float().norm() + 1e-9)).item()
    assert rel(y, yr[rank * rows:(rank + 1) * rows]) < 2e-2
    assert rel(xl.grad, Xr.grad[rank * rows:(rank + 1) * rows]) < 3e-2
    assert rel(w1.grad, W1r.grad[rank * il:(rank + 1) * il]) < 3e-2
    assert rel(bb1.grad, b1r.grad[rank * il:(rank + 1) * il]) < 3e-2
    assert rel(w2.grad, W2r.grad[:, rank * il:(rank + 1) * il]) < 3e-2
    if rank == 0:
        assert rel(bb2.grad, b2r.grad) < 3e-2


def test_sequence_parallel_mlp_fused():
    run_distributed(_sp_mlp, 2, backend="nccl")


def _watchdog_fault(rank, world, marker):
    """Fault injection: rank 1 silently drops its all-reduce; rank 0's in-kernel flag wait must hit
    the watchdog and surface as a CUDA error instead of hanging the GPU."""
    import os
    import time
    from quintnet_b200.ops import symm
    dev = torch.device("cuda", rank)
    t = symm.alloc(1 << 16, torch.float32, dev, dist.group.WORLD)
    t.fill_(1.0)
    symm.set_flag_watchdog(0.25)
    torch.cuda.synchronize()
    dist.barrier()
    os.environ["QUINTNET_FAULT"] = "drop_collective:1"
    symm.all_reduce_(symm.lookup(t), t, 1.0)
    if rank == 0:
        t0 = time.time()
        try:
            torch.cuda.synchronize()
            outcome = "no-error"
        except RuntimeError as e:          # trap -> launch failure on the survivor
            outcome = f"fired after {time.time() - t0:.2f}s: {str(e).splitlines()[0]}"
        with open(marker, "w") as f:
            f.write(outcome)
    else:
        time.sleep(4.0)                    # keep the peer mapping alive while rank 0 spins
    os._exit(0)                            # the CUDA context of rank 0 is gone; skip the collective teardown


def test_flag_watchdog_fires_on_dropped_collective(tmp_path):
    marker = str(tmp_path / "watchdog.txt")
    run_distributed(_watchdog_fault, 2, marker, backend="nccl", timeout=120)
    with open(marker) as f:
        outcome = f.read()
    print(outcome)
    assert outcome.startswith("fired"), outcome


# --------------------------------------------------------------------------------------------------------
# 4-rank variants of the in-kernel collectives (group sizes beyond the 2-rank axes of the 2x2x2 mesh).  Written
# without hardware access in round 1 -> opt-in until they have run once (QN_RUN_UNVALIDATED=1).
# --------------------------------------------------------------------------------------------------------
_unvalidated = pytest.mark.gpu          # (were opt-in in round 1; they run by default now)


@_unvalidated
@pytest.mark.multigpu(4)
@pytest.mark.parametrize("fn", ["_allreduce", "_oneshot", "_zero1_native", "_tp_fused", "_sp_mlp"])
def test_in_kernel_collectives_world4(fn):
    run_distributed(globals()[fn], 4, backend="nccl")
