"""Coordinators compose the parallelism primitives in the fixed order
TP (innermost) → PP → DP (outermost) — reference ``coordinators/`` (718 LoC,
``main_coordinator.py:36-60`` + 7 concrete classes).  Here the composition
logic lives once in :func:`compose`; the concrete coordinators only choose
which axes are active.  Fixes: DP always wraps on the *dp axis group* (the
reference's ``dp``/``dp_tp`` coordinators used the world group);
missing axes are size 1 (D8).
"""
from __future__ import annotations

import os
from abc import ABC, abstractmethod
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from ..parallel.data_parallel import BucketConfig, DataParallel, DistributedConfig
from ..parallel.pipeline_parallel import PipelineParallelWrapper
from ..parallel.tensor_parallel import apply_tensor_parallel

_DTYPES = {"float32": torch.float32, "fp32": torch.float32, "bfloat16": torch.bfloat16,
           "bf16": torch.bfloat16, "float16": torch.float16, "fp16": torch.float16}


def resolve_dtype(config: dict) -> torch.dtype:
    v = config.get("dtype", "float32")
    return v if isinstance(v, torch.dtype) else _DTYPES[str(v).lower()]


def resolve_main_grad_dtype(config: dict) -> Optional[torch.dtype]:
    """dtype of the gradient buckets when it differs from the parameter dtype.  Low-precision training (bf16 / fp16
    parameters) accumulates micro-batch gradients and reduces them over DP in fp32 main-grad buckets by default
    (``main_grad_dtype: float32``; set ``main_grad_dtype: param`` to accumulate in the parameter dtype)."""
    if resolve_dtype(config) == torch.float32:
        return None
    v = config.get("main_grad_dtype", "float32")
    if v is None or str(v).lower() in ("param", "none", "same"):
        return None
    dt = v if isinstance(v, torch.dtype) else _DTYPES[str(v).lower()]
    return None if dt == resolve_dtype(config) else dt


def resolve_device(pg_manager, config: Optional[dict] = None) -> torch.device:
    dt = (config or {}).get("device_type", getattr(pg_manager, "device_type", "cuda"))
    if dt == "cuda" and torch.cuda.is_available():
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    return torch.device("cpu")


def wrap_data_parallel(model: nn.Module, pg_manager, config: dict) -> nn.Module:
    dp_size = pg_manager.axis_size("dp")
    zero = int(config.get("zero_stage", 0))
    if dp_size == 1 and zero == 0 and not config.get("force_dp_wrapper", False):
        return model
    dist_cfg = DistributedConfig(rank=pg_manager.axis_rank("dp"), world_size=dp_size,
                                 process_group=pg_manager.get_group("dp"))
    bcfg = BucketConfig(capacity_mb=int(config.get("bucket_cap_mb", 25)),
                        flatten_parameters=zero >= 1 or bool(config.get("flatten_parameters", False)),
                        grad_dtype=resolve_main_grad_dtype(config))
    return DataParallel(model, dist_cfg, bcfg, reduce_mode="defer" if zero >= 1 else "all_reduce")


def make_tied_group(pg_manager):
    """{first, last} stage sub-group of this rank's pp fibre (created collectively)."""
    pp = pg_manager.axis_size("pp")
    if pp <= 1:
        return None
    if pp == 2:
        return pg_manager.get_group("pp")
    mine = None
    for fibre in pg_manager.device_mesh.get_all_fibres("pp"):
        g = dist.new_group(ranks=[fibre[0], fibre[-1]], backend=pg_manager.device_mesh.backend)
        if pg_manager.rank in (fibre[0], fibre[-1]):
            mine = g
    return mine


def compose(model: Optional[nn.Module], pg_manager, config: dict, *, use_tp: bool, use_pp: bool,
            use_dp: bool, stage_module: Optional[nn.Module] = None, stage_chunks=None) -> nn.Module:
    """``stage_chunks``: the V model chunks of this rank for the interleaved (virtual-stage) pipeline."""
    device = resolve_device(pg_manager, config)
    dtype = resolve_dtype(config)
    if use_pp and pg_manager.axis_size("pp") > 1 and bool(config.get("pp_direction_channels", True)):
        # one communicator per pipeline direction: pre-posted activation receives overlap the running backward
        from ..core.communication import register_pipeline_channels
        fwd_g, bwd_g = pg_manager.device_mesh.get_channel_groups("pp", 2)
        if fwd_g is not None and bwd_g is not None:
            register_pipeline_channels(pg_manager.get_group("pp"), fwd_g, bwd_g)
    if stage_chunks is not None:
        from ..parallel.pipeline_parallel.interleaved import VirtualPipelineWrapper
        model = VirtualPipelineWrapper(stage_chunks, pg_manager.device_mesh, pg_manager.axis_rank("pp"),
                                       pg_manager.get_group("pp"), pg_manager.axis_size("pp"), device)
        return wrap_data_parallel(model, pg_manager, config) if use_dp else model
    if model is not None:
        model = model.to(device=device, dtype=dtype) if dtype != torch.float32 else model.to(device)
    if use_tp and model is not None and pg_manager.axis_size("tp") > 1:
        model = apply_tensor_parallel(
            model, pg_manager.axis_size("tp"), pg_manager.axis_rank("tp"), pg_manager.get_group("tp"),
            device, gather_output=True, sync_gradients=True,
            method_of_parallelism=config.get("tp_method", "column"))
    if use_pp and (pg_manager.axis_size("pp") > 1 or stage_module is not None
                   or config.get("force_pp_wrapper", False)):
        model = PipelineParallelWrapper(model, pg_manager.device_mesh, pg_manager.axis_rank("pp"),
                                        pg_manager.get_group("pp"), pg_manager.axis_size("pp"), device,
                                        stage_module=stage_module)
    elif stage_module is not None:
        model = stage_module.to(device)
    if use_dp:
        model = wrap_data_parallel(model, pg_manager, config)
    return model


class BaseCoordinator(ABC):
    def __init__(self, model: Optional[nn.Module], **kwargs):
        self.model = model
        for k, v in kwargs.items():
            setattr(self, k, v)

    @abstractmethod
    def parallelize(self) -> nn.Module: ...
