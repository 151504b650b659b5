from .main_coordinator import BaseCoordinator, compose, wrap_data_parallel, resolve_device, resolve_dtype
from .concrete import (DataParallelCoordinator, TensorParallelCoordinator, PipelineParallelCoordinator,
                       DPTCoordinator, DPPCoordinator, TPPCoordinator, Hybrid3DCoordinator)

__all__ = ["BaseCoordinator", "DataParallelCoordinator", "TensorParallelCoordinator",
           "PipelineParallelCoordinator", "DPTCoordinator", "DPPCoordinator", "TPPCoordinator",
           "Hybrid3DCoordinator", "compose", "wrap_data_parallel", "resolve_device", "resolve_dtype"]
