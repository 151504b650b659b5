"""Import-compatibility shim: ``import QuintNet`` gives the reference's public surface
(``__init__.py:17-54`` + the README quick-start's ``Trainer``) backed by :mod:`quintnet_b200`.

    from QuintNet import init_process_groups, get_strategy, Trainer
    from QuintNet.core import load_config, All_Reduce
    from QuintNet.parallelism.tensor_parallel.layers import ColumnParallelLinear
    from QuintNet.utils.GPT2.gpt2_stage import GPT2Stage
    from QuintNet.GPT2_Trainer import GPT2Trainer

Every module path of the reference tree resolves: a meta-path finder translates ``QuintNet.<path>`` with the
longest matching prefix of the table below and hands back THE module object of the implementation (no second copy
of any module is ever executed, so classes keep their identity across both names).
"""
import importlib
import importlib.abc
import importlib.util
import sys

from quintnet_b200 import *  # noqa: F401,F403
from quintnet_b200 import __all__ as _all, __version__  # noqa: F401

# reference path (relative to the package) -> implementation module; longest prefix wins, the rest of the dotted path
# is appended unchanged (e.g. parallelism.data_parallel.components.bucket -> …parallel.data_parallel.components.bucket)
_PREFIXES = {
    "core": "quintnet_b200.core",
    "parallelism": "quintnet_b200.parallel",
    "strategy": "quintnet_b200.strategy",
    "coordinators": "quintnet_b200.coordinators",
    "optimizers": "quintnet_b200.optim",
    "ops": "quintnet_b200.ops",
    "models": "quintnet_b200.models",
    "tools": "quintnet_b200.tools",
    "utils": "quintnet_b200.utils",
    "utils.GPT2": "quintnet_b200.models.gpt2",
    "trainer": "quintnet_b200.trainer",
    "GPT2_Trainer": "quintnet_b200.gpt2_trainer",
    "merge_checkpoints": "quintnet_b200.tools.merge_checkpoints",
}
# reference files whose content lives in a differently named module here
_EXACT = {
    "utils.model": "quintnet_b200.models.vit",
    "utils.Dataloader": "quintnet_b200.utils.data",
    "utils.logging": "quintnet_b200.utils.logger",              # setup_logger / log_rank_0
    "optimizers.zero": "quintnet_b200.optim.zero",
    "optimizers.distributed_adamw": "quintnet_b200.optim.distributed_adamw",
    **{f"strategy.{n}": "quintnet_b200.strategy.strategies" for n in (
        "data_parallel_strategy", "tensor_parallel_strategy", "pipeline_parallel_strategy", "dp_tp_strategy",
        "dp_pp_strategy", "tp_pp_strategy", "hybrid_3d_strategy")},
    **{f"coordinators.{n}": "quintnet_b200.coordinators.concrete" for n in (
        "data_parallel_coordinator", "tensor_parallel_coordinator", "pipeline_parallel_coordinator",
        "dp_tp_coordinator", "dp_pp_coordinator", "tp_pp_coordinator", "hybrid_3d_coordinator")},
}


def _translate(rel: str):
    if rel in _EXACT:
        return _EXACT[rel]
    parts = rel.split(".")
    for cut in range(len(parts), 0, -1):
        head = ".".join(parts[:cut])
        if head in _PREFIXES:
            return ".".join([_PREFIXES[head]] + parts[cut:])
    return None


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target: str):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)      # the real module object

    def exec_module(self, module):                       # already executed under its own name
        return None


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(__name__ + "."):
            return None
        real = _translate(fullname[len(__name__) + 1:])
        if real is None:
            return None
        try:
            importlib.import_module(real)
        except ModuleNotFoundError:
            return None
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(real))
        return spec


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())

for _name in ("core", "parallelism", "strategy", "coordinators", "optimizers", "utils", "trainer", "GPT2_Trainer",
              "merge_checkpoints"):
    globals()[_name] = importlib.import_module(f"{__name__}.{_name}")
__all__ = list(_all)
