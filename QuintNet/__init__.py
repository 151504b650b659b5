"""Import-compatibility shim: ``import QuintNet`` gives the reference's public surface
(``__init__.py:17-54`` + the README quick-start's ``Trainer``) backed by :mod:`quintnet_b200`.

    from QuintNet import init_process_groups, get_strategy, Trainer
    from QuintNet.core import load_config, All_Reduce
    from QuintNet.parallelism.tensor_parallel import ColumnParallelLinear
    from QuintNet.utils.GPT2 import GPT2Config, GPT2Stage
"""
import importlib
import sys

import quintnet_b200 as _q
from quintnet_b200 import *  # noqa: F401,F403
from quintnet_b200 import __all__ as _all, __version__  # noqa: F401

_ALIASES = {
    "core": "quintnet_b200.core",
    "core.mesh": "quintnet_b200.core.mesh",
    "core.process_groups": "quintnet_b200.core.process_groups",
    "core.communication": "quintnet_b200.core.communication",
    "core.config": "quintnet_b200.core.config",
    "core.distributed": "quintnet_b200.core.distributed",
    "core.distributed_loading": "quintnet_b200.core.distributed_loading",
    "parallelism": "quintnet_b200.parallel",
    "parallelism.data_parallel": "quintnet_b200.parallel.data_parallel",
    "parallelism.tensor_parallel": "quintnet_b200.parallel.tensor_parallel",
    "parallelism.pipeline_parallel": "quintnet_b200.parallel.pipeline_parallel",
    "strategy": "quintnet_b200.strategy",
    "coordinators": "quintnet_b200.coordinators",
    "optimizers": "quintnet_b200.optim",
    "utils": "quintnet_b200.utils",
    "utils.GPT2": "quintnet_b200.models.gpt2",
    "utils.model": "quintnet_b200.models.vit",
    "utils.Dataloader": "quintnet_b200.utils.data",
    "utils.metrics": "quintnet_b200.utils.metrics",
    "utils.checkpoint": "quintnet_b200.utils.checkpoint",
    "trainer": "quintnet_b200.trainer",
    "GPT2_Trainer": "quintnet_b200.gpt2_trainer",
    "merge_checkpoints": "quintnet_b200.tools.merge_checkpoints",
}
for _name, _target in _ALIASES.items():
    _mod = importlib.import_module(_target)
    sys.modules[f"{__name__}.{_name}"] = _mod
    if "." not in _name:
        globals()[_name] = _mod
__all__ = list(_all)
