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
