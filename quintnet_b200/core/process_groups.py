"""``ProcessGroupManager`` / ``init_process_groups`` — the user entry point.

API parity: reference ``core/process_groups.py:42-180``.  Differences: cpu
(gloo) meshes are accepted, 1-D/2-D meshes work, axis sizes are looked up by
*name* instead of position.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch.distributed as dist

from .mesh import DeviceMesh


class ProcessGroupManager:
    def __init__(
        self,
        device_type: str = "cuda",
        mesh_dim: Sequence[int] = (2, 2, 2),
        mesh_name: Sequence[str] = ("dp", "pp", "tp"),
        backend: Optional[str] = None,
    ):
        if device_type not in ("cuda", "cpu"):
            raise ValueError(f"device_type must be 'cuda' or 'cpu', got {device_type!r}")
        if len(mesh_dim) != len(mesh_name):
            raise ValueError(
                f"mesh_dim {tuple(mesh_dim)} and mesh_name {tuple(mesh_name)} differ in length"
            )
        if any(int(d) < 1 for d in mesh_dim):
            raise ValueError(f"mesh dims must be >= 1, got {tuple(mesh_dim)}")
        self.device_type = device_type
        self.mesh_dim = tuple(int(d) for d in mesh_dim)
        self.mesh_name = tuple(mesh_name)
        self.device_mesh = DeviceMesh(
            mesh_dim=self.mesh_dim, mesh_name=self.mesh_name, device_type=device_type, backend=backend
        )
        self.mesh = self.device_mesh.mesh

    # -- reference-compatible accessors
    def get_group(self, dim_name: str):
        return self.device_mesh.get_group(dim_name)

    def get_all_groups(self) -> Dict[str, Optional[dist.ProcessGroup]]:
        return self.device_mesh.get_all_groups()

    def get_coordinates_tensor_search(self, rank: Optional[int] = None) -> List[int]:
        return self.device_mesh.get_coordinates(rank)

    # -- name-based helpers used throughout the new code
    def axis_size(self, name: str) -> int:
        return self.device_mesh.axis_size(name)

    def axis_rank(self, name: str, rank: Optional[int] = None) -> int:
        return self.device_mesh.axis_rank(name, rank)

    def get_group_ranks(self, name: str) -> List[int]:
        return self.device_mesh.get_group_ranks(name)

    @property
    def rank(self) -> int:
        return self.device_mesh.rank

    @property
    def world_size(self) -> int:
        return self.device_mesh.world_size

    def print_mesh_info(self) -> None:
        if self.rank != 0:
            return
        print("=" * 60)
        print(f"Device mesh {dict(zip(self.mesh_name, self.mesh_dim))} "
              f"({self.world_size} ranks, backend {self.device_mesh.backend})")
        header = " rank | " + " | ".join(f"{n:>3}" for n in self.mesh_name)
        print(header)
        print("-" * len(header))
        for r in range(self.world_size):
            coords = self.device_mesh.get_coordinates(r)
            print(f" {r:4d} | " + " | ".join(f"{c:3d}" for c in coords))
        for name in self.mesh_name:
            print(f" {name} groups: {self.device_mesh.get_all_fibres(name)}")
        print("=" * 60)


def init_process_groups(
    device_type: str = "cuda",
    mesh_dim: Sequence[int] = (2, 2, 2),
    mesh_name: Sequence[str] = ("dp", "pp", "tp"),
    backend: Optional[str] = None,
) -> ProcessGroupManager:
    return ProcessGroupManager(device_type, mesh_dim, mesh_name, backend)
