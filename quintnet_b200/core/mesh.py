"""N-D named device mesh and per-axis process groups.

Capability parity with the reference's ``MeshGenerator`` (reference
``core/mesh.py:124-294``) but redesigned:

* any number of mesh axes >= 1 (the reference crashes on 1-D/2-D meshes,
  SURVEY §8-D8); axes that are not named are treated as size 1;
* *every* rank creates *every* sub-group in the same canonical order (the
  reference only creates the row it belongs to);
* backend follows the device type: ``nccl`` for cuda, ``gloo`` for cpu, so
  the whole framework can be plumbing-tested on CPU;
* a single-process mesh (world size 1) works without ``torch.distributed``
  groups at all — every axis group is then ``None`` and collectives short-cut.
"""
from __future__ import annotations

import datetime
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

_KNOWN_AXES = ("dp", "tp", "pp", "cp")


def _default_backend(device_type: str) -> str:
    return "nccl" if device_type == "cuda" else "gloo"


class DeviceMesh:
    """Named N-D mesh of global ranks.

    ``mesh`` is ``arange(prod(dims)).view(dims)``; the sub-groups of axis ``a``
    are the 1-D fibres of the mesh along ``a``.  ``get_group(name)`` returns the
    fibre that contains the calling rank.
    """

    def __init__(
        self,
        mesh: Optional[torch.Tensor] = None,
        device_type: str = "cuda",
        mesh_dim: Optional[Sequence[int]] = None,
        mesh_name: Sequence[str] = ("dp", "pp", "tp"),        # the reference's default axis order
        is_backend_initialised: bool = False,
        backend: Optional[str] = None,
        timeout_seconds: int = 1800,
    ):
        # positional order and ``is_backend_initialised`` as in the reference (``core/mesh.py:132``); the flag is
        # informational here — the default process group is created only if it does not exist yet
        if mesh is None:
            if mesh_dim is None:
                raise ValueError("either mesh or mesh_dim must be given")
            n = 1
            for d in mesh_dim:
                n *= int(d)
            mesh = torch.arange(n).view(*[int(d) for d in mesh_dim])
        if mesh.dim() != len(mesh_name):
            raise ValueError(
                f"mesh has {mesh.dim()} dims but {len(mesh_name)} axis names were given"
            )
        if len(set(mesh_name)) != len(mesh_name):
            raise ValueError(f"duplicate axis names in {mesh_name}")
        self.mesh = mesh.clone().cpu()
        self.mesh_dim: Tuple[int, ...] = tuple(int(s) for s in self.mesh.shape)
        self.mesh_name: Tuple[str, ...] = tuple(str(n) for n in mesh_name)
        self.device_type = device_type
        self.backend = backend or _default_backend(device_type)
        self.timeout = datetime.timedelta(seconds=timeout_seconds)
        self.world_size = int(self.mesh.numel())

        self._setup_default_group_and_device()
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if self.rank >= self.world_size:
            raise ValueError(f"rank {self.rank} is outside a mesh of {self.world_size} ranks")

        self._groups: Dict[str, Optional[dist.ProcessGroup]] = {}
        self._group_ranks: Dict[str, List[int]] = {}
        self._all_fibres: Dict[str, List[List[int]]] = {}
        self._init_process_groups()

    # ------------------------------------------------------------------ setup
    def _setup_default_group_and_device(self) -> None:
        if self.device_type == "cuda" and torch.cuda.is_available():
            local_rank = int(os.environ.get("LOCAL_RANK", 0))
            torch.cuda.set_device(local_rank)
        if self.world_size > 1 and not dist.is_initialized():
            kwargs = {}
            if self.device_type == "cuda" and torch.cuda.is_available():
                kwargs["device_id"] = torch.device("cuda", torch.cuda.current_device())
            dist.init_process_group(backend=self.backend, timeout=self.timeout, **kwargs)
        if dist.is_initialized() and dist.get_world_size() != self.world_size:
            raise ValueError(
                f"mesh {self.mesh_dim} needs {self.world_size} ranks but the job has "
                f"{dist.get_world_size()}"
            )

    def _init_process_groups(self) -> None:
        for axis, name in enumerate(self.mesh_name):
            fibres = self.mesh.movedim(axis, -1).reshape(-1, self.mesh_dim[axis]).tolist()
            self._all_fibres[name] = fibres
            mine: Optional[List[int]] = None
            my_group = None
            for ranks in fibres:
                group = None
                if self.world_size > 1 and len(ranks) > 1:
                    # canonical creation: all ranks call new_group for all fibres
                    group = dist.new_group(ranks=ranks, backend=self.backend, timeout=self.timeout)
                if self.rank in ranks:
                    mine, my_group = ranks, group
            assert mine is not None
            self._groups[name] = my_group
            self._group_ranks[name] = mine

    def get_channel_groups(self, name: str, n: int = 2):
        """``n`` extra process groups over this rank's fibre of axis ``name`` (same ranks, separate communicators) —
        e.g. one per pipeline direction so that pre-posted receives do not queue behind sends.  COLLECTIVE: every rank of
        the job must call it (all fibres are created canonically, like the axis groups).  Cached."""
        key = f"{name}#channels{n}"
        if key in self._groups:
            return self._groups[key]
        mine = [None] * n
        if name in self.mesh_name and self.world_size > 1:
            for ranks in self._all_fibres[name]:
                for c in range(n):
                    g = dist.new_group(ranks=ranks, backend=self.backend, timeout=self.timeout) if len(ranks) > 1 else None
                    if self.rank in ranks:
                        mine[c] = g
        self._groups[key] = tuple(mine)
        return self._groups[key]

    # ---------------------------------------------------------------- queries
    def axis_size(self, name: str) -> int:
        """Size of a named axis; an axis that is not part of the mesh has size 1."""
        return self.mesh_dim[self.mesh_name.index(name)] if name in self.mesh_name else 1

    def axis_rank(self, name: str, rank: Optional[int] = None) -> int:
        if name not in self.mesh_name:
            return 0
        return self.get_coordinates(rank)[self.mesh_name.index(name)]

    @property
    def groups(self) -> Dict[str, Optional[dist.ProcessGroup]]:
        """axis name -> this rank's process group (the reference's ``self.groups`` dict)."""
        return dict(self._groups)

    def get_group(self, dim_name: str) -> Optional[dist.ProcessGroup]:
        name = dim_name
        if name not in self.mesh_name:
            if name in _KNOWN_AXES:
                return None
            raise KeyError(f"unknown mesh axis '{name}' (have {self.mesh_name})")
        return self._groups[name]

    def get_group_ranks(self, name: str) -> List[int]:
        if name not in self.mesh_name:
            return [self.rank]
        return list(self._group_ranks[name])

    def get_all_fibres(self, name: str) -> List[List[int]]:
        return [list(f) for f in self._all_fibres.get(name, [[r] for r in range(self.world_size)])]

    def get_all_groups(self) -> Dict[str, Optional[dist.ProcessGroup]]:
        return dict(self._groups)

    def get_coordinates(self, rank: Optional[int] = None) -> List[int]:
        rank = self.rank if rank is None else int(rank)
        idx = (self.mesh == rank).nonzero()
        if idx.numel() == 0:
            raise ValueError(f"rank {rank} not in mesh")
        return idx[0].tolist()

    # reference-compatible spelling (core/mesh.py:268)
    def get_coordinates_tensor_search(self, rank: Optional[int] = None) -> List[int]:
        return self.get_coordinates(rank)

    @property
    def dp_size(self) -> int:
        return self.axis_size("dp")

    @property
    def tp_size(self) -> int:
        return self.axis_size("tp")

    @property
    def pp_size(self) -> int:
        return self.axis_size("pp")

    def __repr__(self) -> str:
        dims = ", ".join(f"{n}={d}" for n, d in zip(self.mesh_name, self.mesh_dim))
        return f"DeviceMesh({dims}; backend={self.backend}; rank={self.rank}/{self.world_size})"


# reference name
MeshGenerator = DeviceMesh
