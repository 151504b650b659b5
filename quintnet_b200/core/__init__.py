from .mesh import DeviceMesh, MeshGenerator
from .process_groups import ProcessGroupManager, init_process_groups
from .communication import (Send, Recv, All_Gather, All_Reduce, ReduceScatter, Copy_To_Parallel,
                            Scatter_To_Parallel, pipeline_communicate,
                            bidirectional_pipeline_communicate, drain_pipeline_sends)
from .config import load_config, merge_configs, ParallelismConfig, TrainingConfig, axis_size
from .distributed import (setup_distributed, cleanup_distributed, get_rank, get_world_size,
                          is_main_process)
from .distributed_loading import (load_gpt2_distributed, random_gpt2_state_dict,
                                  save_random_hf_checkpoint)

__all__ = ["DeviceMesh", "MeshGenerator", "ProcessGroupManager", "init_process_groups", "Send", "Recv",
           "All_Gather", "All_Reduce", "ReduceScatter", "Copy_To_Parallel", "Scatter_To_Parallel",
           "pipeline_communicate", "bidirectional_pipeline_communicate", "drain_pipeline_sends",
           "load_config", "merge_configs", "ParallelismConfig", "TrainingConfig", "axis_size",
           "setup_distributed", "cleanup_distributed", "get_rank", "get_world_size", "is_main_process",
           "load_gpt2_distributed", "random_gpt2_state_dict", "save_random_hf_checkpoint"]
