"""quintnet_b200 — a Blackwell-native 3-D parallel training framework with the
capabilities and public API of QuintNet (reference ``__init__.py:17-54``),
rebuilt B200-first: PyTorch for the framework layer, hand-written sm_100a
kernels (tcgen05/TMEM GEMMs, fused elementwise/normalisation/optimizer kernels,
in-kernel NVLink collectives) for the hot paths.
"""
__version__ = "0.1.0"

from .core import (DeviceMesh, MeshGenerator, ProcessGroupManager, init_process_groups,
                   setup_distributed, cleanup_distributed, load_config, merge_configs)
from .parallel import (DataParallel, TensorParallel, apply_tensor_parallel, ColumnParallelLinear,
                       RowParallelLinear, VocabParallelEmbedding, PipelineParallelWrapper,
                       PipelineTrainer, PipelineDataLoader)
from .strategy import (get_strategy, BaseStrategy, DataParallelStrategy, TensorParallelStrategy,
                       PipelineParallelStrategy, DataTensorParallelStrategy,
                       DataPipelineParallelStrategy, TensorPipelineParallelStrategy, Hybrid3DStrategy)
from .optim import FusedAdamW, DistributedAdamW
from .trainer import Trainer
from .gpt2_trainer import GPT2Trainer

__all__ = [
    "DeviceMesh", "MeshGenerator", "ProcessGroupManager", "init_process_groups", "setup_distributed",
    "cleanup_distributed", "load_config", "merge_configs", "DataParallel", "TensorParallel",
    "apply_tensor_parallel", "ColumnParallelLinear", "RowParallelLinear", "VocabParallelEmbedding",
    "PipelineParallelWrapper", "PipelineTrainer", "PipelineDataLoader", "get_strategy", "BaseStrategy",
    "DataParallelStrategy", "TensorParallelStrategy", "PipelineParallelStrategy",
    "DataTensorParallelStrategy", "DataPipelineParallelStrategy", "TensorPipelineParallelStrategy",
    "Hybrid3DStrategy", "FusedAdamW", "DistributedAdamW", "Trainer", "GPT2Trainer",
]
