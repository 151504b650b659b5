"""3-D composition on CPU / gloo (reference ``tests/test_hybrid.py``: ``test_3d_initialization``, ``test_coordinator`` —
8 GPUs there)."""
import torch
import torch.distributed as dist

from tests._dist import run_distributed
from quintnet_b200 import DataParallel, PipelineParallelWrapper, get_strategy, init_process_groups
from quintnet_b200.coordinators import (DataParallelCoordinator, DPPCoordinator, DPTCoordinator, Hybrid3DCoordinator,
                                        PipelineParallelCoordinator, TensorParallelCoordinator, TPPCoordinator)
from quintnet_b200.models import Model
from quintnet_b200.parallel.tensor_parallel import ColumnParallelLinear, RowParallelLinear

CFG = {"device_type": "cpu", "mesh_dim": [2, 2, 2], "mesh_name": ["dp", "tp", "pp"], "strategy_name": "3d",
       "batch_size": 16, "grad_acc_steps": 2, "tp_method": "paired"}


def _init_3d(rank, world):
    pg = init_process_groups("cpu", CFG["mesh_dim"], CFG["mesh_name"])
    coords = pg.get_coordinates_tensor_search(rank)
    assert coords == [rank // 4, (rank // 2) % 2, rank % 2]
    for axis in ("dp", "tp", "pp"):
        assert dist.get_world_size(pg.get_group(axis)) == 2 and pg.axis_rank(axis) == coords[CFG["mesh_name"].index(axis)]
    torch.manual_seed(0)
    model = get_strategy("3d", pg, CFG).apply(Model(hidden_dim=32, depth=4, n_heads=4))
    # composition order TP -> PP -> DP: DataParallel( PipelineParallelWrapper( stage with TP layers ) )
    assert isinstance(model, DataParallel) and isinstance(model.model, PipelineParallelWrapper)
    stage = model.model.local_module
    kinds = {type(m) for m in stage.modules()}
    assert ColumnParallelLinear in kinds and RowParallelLinear in kinds
    n_blocks = sum(1 for m in stage if m.__class__.__name__ == "TransformerBlock")
    assert n_blocks == 2                                                    # 4 blocks over pp = 2
    # replicas agree after the wrap-time broadcast; tp shards differ
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    peers = [torch.zeros_like(flat) for _ in range(2)]
    dist.all_gather(peers, flat, group=pg.get_group("dp"))
    assert torch.equal(peers[0], peers[1])


def test_3d_initialization():
    run_distributed(_init_3d, 8)


def _coordinators(rank, world):
    pg = init_process_groups("cpu", [2, 2], ["tp", "pp"])
    cfg = dict(CFG, mesh_dim=[2, 2], mesh_name=["tp", "pp"])
    mk = lambda: Model(hidden_dim=32, depth=4, n_heads=4)
    torch.manual_seed(1)
    m = TPPCoordinator(mk(), pg_manager=pg, config=cfg).parallelize()
    assert isinstance(m, PipelineParallelWrapper) and any(isinstance(x, ColumnParallelLinear) for x in m.modules())
    m = TensorParallelCoordinator(mk(), pg_manager=pg, config=cfg).parallelize()
    assert not isinstance(m, PipelineParallelWrapper) and any(isinstance(x, RowParallelLinear) for x in m.modules())
    m = PipelineParallelCoordinator(mk(), pg_manager=pg, config=cfg).parallelize()
    assert isinstance(m, PipelineParallelWrapper) and not any(isinstance(x, ColumnParallelLinear) for x in m.modules())
    # dp-flavoured coordinators on a mesh without a dp axis degrade to their other axes (axis size 1)
    for cls in (DataParallelCoordinator, DPTCoordinator, DPPCoordinator, Hybrid3DCoordinator):
        out = cls(mk(), pg_manager=pg, config=cfg).parallelize()
        assert sum(p.numel() for p in out.parameters()) > 0


def test_coordinator():
    run_distributed(_coordinators, 4)
