"""Mesh / process-group plumbing on CPU/gloo (reference tests/test_mesh.py needs
4–8 real GPUs and crashes on 2-D meshes)."""
import pytest
import torch
import torch.distributed as dist

from tests._dist import run_distributed
from quintnet_b200 import init_process_groups


def _mesh_2x2(rank, world):
    pg = init_process_groups("cpu", [2, 2], ["dp", "tp"])
    assert pg.axis_size("dp") == 2 and pg.axis_size("tp") == 2 and pg.axis_size("pp") == 1
    assert pg.get_group("pp") is None
    coords = pg.get_coordinates_tensor_search(rank)
    assert coords == [rank // 2, rank % 2]
    assert dist.get_world_size(pg.get_group("tp")) == 2
    assert sorted(pg.get_group_ranks("tp")) == [2 * (rank // 2), 2 * (rank // 2) + 1]
    assert sorted(pg.get_group_ranks("dp")) == [rank % 2, rank % 2 + 2]
    # group really works: sum of ranks along tp
    t = torch.tensor([float(rank)])
    dist.all_reduce(t, group=pg.get_group("tp"))
    assert t.item() == sum(pg.get_group_ranks("tp"))
    pg.print_mesh_info()


def _mesh_2x2x2(rank, world):
    pg = init_process_groups("cpu", [2, 2, 2], ["dp", "tp", "pp"])
    c = pg.get_coordinates_tensor_search(rank)
    assert c == [rank // 4, (rank // 2) % 2, rank % 2]
    assert pg.device_mesh.dp_size == pg.device_mesh.tp_size == pg.device_mesh.pp_size == 2
    for name in ("dp", "tp", "pp"):
        g = pg.get_group(name)
        assert dist.get_world_size(g) == 2
        assert dist.get_rank(g) == pg.axis_rank(name)
    assert set(pg.get_all_groups()) == {"dp", "tp", "pp"}


def _mesh_1d(rank, world):
    pg = init_process_groups("cpu", [2], ["dp"])
    assert pg.axis_size("dp") == 2 and pg.axis_size("tp") == 1
    assert pg.get_coordinates_tensor_search(rank) == [rank]


def test_mesh_2d():
    run_distributed(_mesh_2x2, 4)


def test_mesh_3d():
    run_distributed(_mesh_2x2x2, 8)


def test_mesh_1d():
    run_distributed(_mesh_1d, 2)


def test_single_process_mesh():
    pg = init_process_groups("cpu", [1, 1, 1], ["dp", "tp", "pp"])
    assert pg.world_size == 1 and pg.get_group("dp") is None


def test_bad_args():
    with pytest.raises(ValueError):
        init_process_groups("tpu", [1], ["dp"])
    with pytest.raises(ValueError):
        init_process_groups("cpu", [1, 1], ["dp"])


def _mesh_4d_with_cp(rank, world):
    """A fourth, arbitrarily named axis (e.g. a context-parallel 'cp' axis reserved for ring attention) composes
    with dp / tp / pp: fibres, coordinates and missing-axis defaults all keep working."""
    pg = init_process_groups("cpu", [2, 1, 2, 2], ["dp", "tp", "pp", "cp"])
    assert pg.axis_size("cp") == 2 and pg.axis_size("tp") == 1 and pg.axis_size("ep") == 1
    coords = pg.get_coordinates_tensor_search(rank)
    assert len(coords) == 4 and coords[3] == pg.axis_rank("cp") == rank % 2
    g = pg.get_group("cp")
    t = torch.tensor([float(rank)])
    dist.all_reduce(t, group=g)
    partner = rank ^ 1
    assert t.item() == rank + partner
    fibres = pg.device_mesh.get_all_fibres("cp")
    assert sorted(map(tuple, fibres)) == [(0, 1), (2, 3), (4, 5), (6, 7)]


def test_mesh_4d_named_axis():
    run_distributed(_mesh_4d_with_cp, 8)
