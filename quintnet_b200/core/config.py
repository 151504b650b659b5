"""YAML config loading (dict-based, as in the reference ``core/config.py:96``)
plus a working ``merge_configs`` and typed views."""
from __future__ import annotations

import copy
import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import yaml


@dataclass
class ParallelismConfig:
    device_type: str = "cuda"
    mesh_dim: List[int] = field(default_factory=lambda: [1, 1, 1])
    mesh_name: List[str] = field(default_factory=lambda: ["dp", "tp", "pp"])
    strategy_name: str = "3d"
    schedule: str = "1f1b"
    timeout_seconds: int = 1800

    def axis(self, name: str) -> int:
        return self.mesh_dim[self.mesh_name.index(name)] if name in self.mesh_name else 1


@dataclass
class TrainingConfig:
    batch_size: int = 32
    learning_rate: float = 1e-4
    num_epochs: int = 1
    grad_acc_steps: int = 1
    max_grad_norm: float = 1.0
    num_workers: int = 0
    weight_decay: float = 0.0
    warmup_steps: int = 0
    dtype: str = "float32"
    optimizer: str = "adam"
    zero_stage: int = 0
    output_dir: str = "./checkpoints"
    log_interval: int = 10
    wandb_project: Optional[str] = None


def _coerce_numbers(obj):
    """YAML reads ``1e-4`` as a string; coerce obvious floats."""
    if isinstance(obj, dict):
        return {k: _coerce_numbers(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_coerce_numbers(v) for v in obj]
    if isinstance(obj, str):
        try:
            if any(c in obj for c in "eE.") and obj.replace(".", "", 1).replace("e", "", 1).replace(
                    "E", "", 1).replace("-", "").replace("+", "").isdigit():
                return float(obj)
        except ValueError:
            pass
    return obj


def load_config(config_path: str) -> Dict[str, Any]:
    if not os.path.exists(config_path):
        raise FileNotFoundError(f"config file not found: {config_path}")
    try:
        with open(config_path, "r") as f:
            cfg = yaml.safe_load(f)
    except yaml.YAMLError as e:
        raise RuntimeError(f"failed to parse config {config_path}: {e}") from e
    if cfg is None:
        cfg = {}
    if not isinstance(cfg, dict):
        raise RuntimeError(f"config {config_path} must be a mapping at top level")
    return _coerce_numbers(cfg)


def merge_configs(base: Dict[str, Any], override: Optional[Dict[str, Any]]) -> Dict[str, Any]:
    """Recursive dict merge; ``override`` wins, nested dicts are merged."""
    out = copy.deepcopy(base)
    for k, v in (override or {}).items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = merge_configs(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


def axis_size(config: Dict[str, Any], name: str) -> int:
    """Size of a named mesh axis in a dict config; missing axis -> 1 (D8)."""
    names = list(config.get("mesh_name", []))
    return int(config["mesh_dim"][names.index(name)]) if name in names else 1


def parallelism_view(config: Dict[str, Any]) -> ParallelismConfig:
    keys = ParallelismConfig.__dataclass_fields__.keys()
    return ParallelismConfig(**{k: config[k] for k in keys if k in config})


def training_view(config: Dict[str, Any]) -> TrainingConfig:
    keys = TrainingConfig.__dataclass_fields__.keys()
    return TrainingConfig(**{k: config[k] for k in keys if k in config})
