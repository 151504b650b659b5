"""``get_strategy(name, pg_manager, config, checkpoint_path=None, is_staged=False)``
— reference ``strategy/__init__.py:52-105`` (case-insensitive, ValueError on unknown)."""
from __future__ import annotations

from typing import Optional

from ..core.process_groups import ProcessGroupManager  # noqa: F401  (re-exported like the reference)
from .base_strategy import BaseStrategy
from .strategies import (DataParallelStrategy, TensorParallelStrategy, PipelineParallelStrategy,
                         DataTensorParallelStrategy, DataPipelineParallelStrategy,
                         TensorPipelineParallelStrategy, Hybrid3DStrategy)

_STRATEGIES = {
    "dp": DataParallelStrategy, "tp": TensorParallelStrategy, "pp": PipelineParallelStrategy,
    "dp_tp": DataTensorParallelStrategy, "dp_pp": DataPipelineParallelStrategy,
    "tp_pp": TensorPipelineParallelStrategy, "3d": Hybrid3DStrategy,
}


def get_strategy(strategy_name: str, pg_manager, config: dict, checkpoint_path: Optional[str] = None,
                 is_staged: bool = False) -> BaseStrategy:
    key = str(strategy_name).lower()
    if key not in _STRATEGIES:
        raise ValueError(f"unknown strategy {strategy_name!r}; available: {sorted(_STRATEGIES)}")
    cls = _STRATEGIES[key]
    if cls is Hybrid3DStrategy:
        return cls(pg_manager, config, checkpoint_path=checkpoint_path, is_staged=is_staged)
    return cls(pg_manager, config)


__all__ = ["get_strategy", "BaseStrategy", "DataParallelStrategy", "TensorParallelStrategy",
           "PipelineParallelStrategy", "DataTensorParallelStrategy", "DataPipelineParallelStrategy",
           "TensorPipelineParallelStrategy", "Hybrid3DStrategy"]
