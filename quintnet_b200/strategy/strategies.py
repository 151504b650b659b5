"""The seven named strategies (reference ``strategy/data_parallel_strategy.py:4``, ``tensor_parallel_strategy.py:4``,
``pipeline_parallel_strategy.py:4``, ``dp_tp_strategy.py:4``, ``dp_pp_strategy.py:4``, ``tp_pp_strategy.py:4``,
``hybrid_3d_strategy.py:13-47``; lookup ``strategy/__init__.py:81-105``).  Each one only names its coordinator;
the composition order TP -> PP -> DP lives in :mod:`quintnet_b200.coordinators`."""
from __future__ import annotations

from typing import Optional

from ..coordinators import (DataParallelCoordinator, DPPCoordinator, DPTCoordinator, Hybrid3DCoordinator,
                            PipelineParallelCoordinator, TensorParallelCoordinator, TPPCoordinator)
from .base_strategy import BaseStrategy


class DataParallelStrategy(BaseStrategy):
    coordinator_cls = DataParallelCoordinator


class TensorParallelStrategy(BaseStrategy):
    coordinator_cls = TensorParallelCoordinator


class PipelineParallelStrategy(BaseStrategy):
    coordinator_cls = PipelineParallelCoordinator


class DataTensorParallelStrategy(BaseStrategy):
    coordinator_cls = DPTCoordinator


class DataPipelineParallelStrategy(BaseStrategy):
    coordinator_cls = DPPCoordinator


class TensorPipelineParallelStrategy(BaseStrategy):
    coordinator_cls = TPPCoordinator


class Hybrid3DStrategy(BaseStrategy):
    coordinator_cls = Hybrid3DCoordinator

    def __init__(self, pg_manager, config: dict, checkpoint_path: Optional[str] = None,
                 is_staged: bool = False):
        super().__init__(pg_manager, config)
        self.checkpoint_path, self.is_staged = checkpoint_path, is_staged

    def coordinator_kwargs(self) -> dict:
        return {"checkpoint_path": self.checkpoint_path, "is_staged": self.is_staged}
