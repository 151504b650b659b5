"""TP layers: forward AND backward equal the dense layer (the reference only
tests forward and its backward is incomplete — SURVEY §8-D2)."""
import torch
import torch.distributed as dist
import torch.nn as nn

from tests._dist import run_distributed
from quintnet_b200 import init_process_groups
from quintnet_b200.core.communication import All_Gather, ReduceScatter, All_Reduce, Copy_To_Parallel
from quintnet_b200.parallel.tensor_parallel import (ColumnParallelLinear, RowParallelLinear,
                                                    VocabParallelEmbedding, apply_tensor_parallel)


def _dense(seed=0, i=16, o=12):
    torch.manual_seed(seed)
    return nn.Linear(i, o)


def _column_row(rank, world):
    pg = init_process_groups("cpu", [world], ["tp"])
    g = pg.get_group("tp")
    dense = _dense()
    torch.manual_seed(7)
    x = torch.randn(4, 5, 16)
    # ---- column, gather_output=True
    xr = x.clone().requires_grad_(True)
    ref = dense(xr)
    ref.pow(2).sum().backward()
    o = 12 // world
    col = ColumnParallelLinear(torch.device("cpu"), g, 16, o, dense.weight.data[rank * o:(rank + 1) * o],
                               dense.bias.data[rank * o:(rank + 1) * o], gather_output=True)
    xc = x.clone().requires_grad_(True)
    y = col(xc)
    y.pow(2).sum().backward()
    assert torch.allclose(y, ref, atol=1e-5)
    assert torch.allclose(xc.grad, xr.grad, atol=1e-4), "input grad must include all TP ranks"
    assert torch.allclose(col.proj.weight.grad, dense.weight.grad[rank * o:(rank + 1) * o], atol=1e-4)
    # ---- row, input_is_parallel=False
    dense.zero_grad()
    xr2 = x.clone().requires_grad_(True)
    ref2 = dense(xr2)
    ref2.pow(2).sum().backward()
    k = 16 // world
    row = RowParallelLinear(torch.device("cpu"), g, k, 12, dense.weight.data[:, rank * k:(rank + 1) * k],
                            dense.bias.data, input_is_parallel=False)
    xc2 = x.clone().requires_grad_(True)
    y2 = row(xc2)
    y2.pow(2).sum().backward()
    assert torch.allclose(y2, ref2, atol=1e-5), "row output incl. bias identical on every rank"
    assert torch.allclose(xc2.grad, xr2.grad, atol=1e-4)
    assert (row.bias is not None) == (rank == 0)


def _column_row_pair_sp(rank, world):
    """Megatron pairing, plain and token-sharded (sequence-parallel) formulations."""
    pg = init_process_groups("cpu", [world], ["tp"])
    g = pg.get_group("tp")
    torch.manual_seed(0)
    fc1, fc2 = nn.Linear(16, 32), nn.Linear(32, 16)
    x = torch.randn(8, 16)
    xr = x.clone().requires_grad_(True)
    ref = fc2(torch.nn.functional.gelu(fc1(xr)))
    ref.pow(2).sum().backward()
    o = 32 // world
    for sp in (False, True):
        col = ColumnParallelLinear(torch.device("cpu"), g, 16, o, fc1.weight.data[rank * o:(rank + 1) * o],
                                   fc1.bias.data[rank * o:(rank + 1) * o], gather_output=False,
                                   sequence_parallel=sp, activation="gelu")
        row = RowParallelLinear(torch.device("cpu"), g, o, 16, fc2.weight.data[:, rank * o:(rank + 1) * o],
                                fc2.bias.data, input_is_parallel=True, sequence_parallel=sp)
        n = 8 // world
        xin = (x[rank * n:(rank + 1) * n] if sp else x).clone().requires_grad_(True)
        y = row(col(xin))
        yref = ref[rank * n:(rank + 1) * n] if sp else ref
        assert torch.allclose(y, yref, atol=1e-5), f"sp={sp}"
        if sp:
            # each rank back-propagates its token shard of the same global loss
            y.pow(2).sum().backward()
            assert torch.allclose(xin.grad, xr.grad[rank * n:(rank + 1) * n], atol=1e-4)
        else:
            y.pow(2).sum().backward()
            assert torch.allclose(xin.grad, xr.grad, atol=1e-4)
        assert torch.allclose(col.proj.weight.grad, fc1.weight.grad[rank * o:(rank + 1) * o], atol=1e-4)
        assert torch.allclose(row.proj.weight.grad, fc2.weight.grad[:, rank * o:(rank + 1) * o], atol=1e-4)


def _vocab_embedding(rank, world):
    pg = init_process_groups("cpu", [world], ["tp"])
    g = pg.get_group("tp")
    torch.manual_seed(0)
    emb = nn.Embedding(20, 8)
    n = 20 // world
    vp = VocabParallelEmbedding(torch.device("cpu"), g, 20, 8, rank * n, (rank + 1) * n,
                                emb.weight.data[rank * n:(rank + 1) * n])
    ids = torch.randint(0, 20, (3, 7))
    assert torch.allclose(vp(ids), emb(ids), atol=1e-6)


def _apply_tp_vit(rank, world):
    from quintnet_b200.models import Model
    pg = init_process_groups("cpu", [world], ["tp"])
    g = pg.get_group("tp")
    for method in ("column", "paired"):
        torch.manual_seed(3)
        ref = Model(hidden_dim=32, depth=2, n_heads=4)
        torch.manual_seed(3)
        tp = Model(hidden_dim=32, depth=2, n_heads=4)
        tp = apply_tensor_parallel(tp, world, rank, g, torch.device("cpu"), method_of_parallelism=method)
        x = torch.randn(4, 1, 28, 28)
        yr = ref(x)
        yt = tp(x)
        assert torch.allclose(yr, yt, atol=1e-4), method
        yr.sum().backward()
        yt.sum().backward()
        assert torch.allclose(ref.embedding.pos_embed.grad, tp.embedding.pos_embed.grad, atol=1e-4), method


def _collectives(rank, world):
    pg = init_process_groups("cpu", [world], ["tp"])
    g = pg.get_group("tp")
    x = (torch.arange(6, dtype=torch.float32).view(2, 3) + 10 * rank).requires_grad_(True)
    y = All_Gather.apply(x, g, "reduce_scatter", 0)
    assert y.shape == (2 * world, 3)
    (y * (rank + 1)).sum().backward()
    assert torch.allclose(x.grad, torch.full_like(x, sum(range(1, world + 1))))
    z = torch.ones(2 * world, 4, requires_grad=True)
    r = ReduceScatter.apply(z * (rank + 1), g, 0)
    assert r.shape == (2, 4) and torch.allclose(r, torch.full_like(r, sum(range(1, world + 1))))
    a = torch.ones(3, requires_grad=True)
    All_Reduce.apply(a * 2, g).sum().backward()
    assert torch.allclose(a.grad, torch.full_like(a, 2.0))
    b = torch.ones(3, requires_grad=True)
    (Copy_To_Parallel.apply(b, g) * (rank + 1)).sum().backward()
    assert torch.allclose(b.grad, torch.full_like(b, float(sum(range(1, world + 1)))))


def test_column_row_forward_backward():
    run_distributed(_column_row, 2)


def test_pair_and_sequence_parallel():
    run_distributed(_column_row_pair_sp, 2)


def test_vocab_parallel_embedding():
    run_distributed(_vocab_embedding, 2)


def test_apply_tensor_parallel_vit():
    run_distributed(_apply_tp_vit, 2)


def test_autograd_collectives():
    run_distributed(_collectives, 4)
