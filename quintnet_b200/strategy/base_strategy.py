"""Strategy API: string name → coordinator (reference ``strategy/``, 306 LoC)."""
from __future__ import annotations

from abc import ABC
from typing import Optional

import torch.nn as nn

from ..coordinators import resolve_device


class BaseStrategy(ABC):
    coordinator_cls = None

    def __init__(self, pg_manager, config: dict):
        self.pg_manager = pg_manager
        self.config = config
        self.device = resolve_device(pg_manager, config)

    def coordinator_kwargs(self) -> dict:
        return {}

    def apply(self, model: Optional[nn.Module] = None) -> nn.Module:
        coord = self.coordinator_cls(model, pg_manager=self.pg_manager, config=self.config,
                                     device=self.device, **self.coordinator_kwargs())
        out = coord.parallelize()
        self.coordinator = coord
        return out
