"""Parallelism primitives: data / tensor / pipeline (reference ``parallelism/``)."""
from .data_parallel import DataParallel
from .tensor_parallel import (apply_tensor_parallel, TensorParallel, ColumnParallelLinear,
                              RowParallelLinear, VocabParallelEmbedding)
from .pipeline_parallel import PipelineParallelWrapper, PipelineTrainer, PipelineDataLoader

__all__ = ["DataParallel", "TensorParallel", "apply_tensor_parallel", "ColumnParallelLinear",
           "RowParallelLinear", "VocabParallelEmbedding", "PipelineParallelWrapper",
           "PipelineTrainer", "PipelineDataLoader"]
