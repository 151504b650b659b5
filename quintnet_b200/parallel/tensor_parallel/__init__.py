from .layers import ColumnParallelLinear, RowParallelLinear, VocabParallelEmbedding
from .model_wrapper import apply_tensor_parallel, TensorParallel
from .utils import (get_tensor_parallel_rank, get_tensor_parallel_world_size,
                    is_tensor_parallel_available, synchronize, print_rank_0, ensure_divisibility)

__all__ = ["ColumnParallelLinear", "RowParallelLinear", "VocabParallelEmbedding",
           "apply_tensor_parallel", "TensorParallel", "get_tensor_parallel_rank",
           "get_tensor_parallel_world_size", "is_tensor_parallel_available", "synchronize",
           "print_rank_0", "ensure_divisibility"]
