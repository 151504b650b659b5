"""DataParallel: bucketed, really-synchronising DDP (the reference's default
config performs no gradient sync at all — SURVEY §8-D1)."""
import copy

import torch
import torch.distributed as dist
import torch.nn as nn

from tests._dist import run_distributed
from quintnet_b200 import init_process_groups, DataParallel
from quintnet_b200.parallel.data_parallel import (BucketConfig, DistributedConfig, ReductionStrategy,
                                                   create_local_ddp, create_distributed_ddp,
                                                   GradientBucket)


def _mlp(seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(10, 16), nn.ReLU(), nn.Linear(16, 16), nn.ReLU(), nn.Linear(16, 5))


def _ddp_equivalence(rank, world):
    pg = init_process_groups("cpu", [world], ["dp"])
    ref = _mlp(100)
    local = _mlp(100 + rank)          # different init per rank: broadcast must fix it
    ddp = DataParallel(local, DistributedConfig(rank, world, pg.get_group("dp")),
                       BucketConfig(capacity_mb=0.001))     # tiny buckets -> several
    assert len(ddp.buckets) > 1
    for p, q in zip(ddp.parameters(), ref.parameters()):
        assert torch.allclose(p, q), "parameters broadcast from dp-rank 0"
    torch.manual_seed(5)
    xs = torch.randn(world, 6, 10)
    ys = torch.randn(world, 6, 5)
    # reference: mean of per-rank grads
    for r in range(world):
        (nn.functional.mse_loss(ref(xs[r]), ys[r]) / world).backward()
    nn.functional.mse_loss(ddp(xs[rank]), ys[rank]).backward()
    for p, q in zip(ddp.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6)
    # second iteration after optimizer-style zero_grad(set_to_none=True)
    for p in ddp.parameters():
        p.grad = None
    ref.zero_grad()
    for r in range(world):
        (nn.functional.mse_loss(ref(xs[r]) * 2, ys[r]) / world).backward()
    nn.functional.mse_loss(ddp(xs[rank]) * 2, ys[rank]).backward()
    for p, q in zip(ddp.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6), "counter reset + re-attached views"


def _ddp_no_sync(rank, world):
    pg = init_process_groups("cpu", [world], ["dp"])
    ref = _mlp(0)
    ddp = DataParallel(_mlp(0), DistributedConfig(rank, world, pg.get_group("dp")))
    torch.manual_seed(9)
    xs = torch.randn(world, 2, 4, 10)
    for r in range(world):
        for mb in range(2):
            (ref(xs[r, mb]).sum() / world).backward()
    with ddp.no_sync():
        ddp(xs[rank, 0]).sum().backward()
        g_local = [p.grad.clone() for p in ddp.parameters()]
    ddp(xs[rank, 1]).sum().backward()
    for p, q in zip(ddp.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-5)
    ddp.zero_grad()
    assert all(float(p.grad.abs().sum()) == 0 for p in ddp.parameters())


def _unused_param(rank, world):
    pg = init_process_groups("cpu", [world], ["dp"])

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(4, 4), nn.Linear(4, 4)

        def forward(self, x):
            return self.a(x)

    torch.manual_seed(0)
    ddp = DataParallel(Net(), DistributedConfig(rank, world, pg.get_group("dp")))
    ddp(torch.ones(2, 4) * (rank + 1)).sum().backward()
    expect = sum(r + 1 for r in range(world)) / world * 2
    assert torch.allclose(ddp.model.a.weight.grad, torch.full((4, 4), expect))


def test_ddp_equivalence():
    run_distributed(_ddp_equivalence, 2)


def test_ddp_no_sync_accumulation():
    run_distributed(_ddp_no_sync, 2)


def test_ddp_unused_parameter():
    run_distributed(_unused_param, 2)


def test_local_ddp_and_bucket_api():
    m = _mlp(0)
    ddp = create_local_ddp(m, BucketConfig(capacity_mb=1))
    ddp(torch.randn(3, 10)).sum().backward()
    info = ddp.get_bucket_info()
    assert info and info[0]["num_params"] == 6
    b = next(iter(ddp.buckets.values()))
    assert isinstance(b, GradientBucket) and b.flatten_gradients().numel() == b.numel
    for p, v in zip(b.parameters, b.grad_views):
        assert p.grad.data_ptr() == v.data_ptr()
    ddp.remove_hooks()
    assert ReductionStrategy.MEAN.value == "mean"


def test_reference_hook_api_names():
    """BucketManager.register_hooks / remove_all_hooks and GradientBucket.has_gradients / register_hooks / remove_hooks
    (the reference's method names) on top of the post-accumulate machinery."""
    import torch.nn as nn
    from quintnet_b200.parallel.data_parallel import BucketConfig, BucketManager
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 4))
    bm = BucketManager(BucketConfig(capacity_mb=1))
    buckets = bm.create_buckets(model)
    ready, seen = [], []
    bm.register_hooks(ready.append)
    b0 = next(iter(buckets.values()))
    b0.register_hooks(lambda p: seen.append(tuple(p.shape)))
    model(torch.randn(3, 8)).sum().backward()
    assert sorted(ready) == sorted(buckets) and b0.has_gradients()
    assert len(seen) == len([p for p in b0.parameters if p.requires_grad])
    bm.remove_all_hooks()
    b0.remove_hooks()
    bm.reset_buckets()
    ready.clear(); seen.clear()
    model(torch.randn(3, 8)).sum().backward()
    assert not ready and not seen
