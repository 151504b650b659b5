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


def test_ddp_nvlink_backend():
    run_distributed(_ddp_native, 2, backend="nccl")


def test_zero1_fused_kernels():
    run_distributed(_zero1_native, 2, backend="nccl")
