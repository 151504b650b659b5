from .core import DataParallel, ReductionStrategy, BucketConfig, DistributedConfig
from .components import GradientBucket, BucketManager, GradientReducer, ParameterBroadcaster
from .backends import DistributedBackend, TorchDistributedBackend, LocalBackend, NVLinkBackend
from .utils import create_local_ddp, create_distributed_ddp

__all__ = ["DataParallel", "ReductionStrategy", "BucketConfig", "DistributedConfig",
           "GradientBucket", "BucketManager", "GradientReducer", "ParameterBroadcaster",
           "DistributedBackend", "TorchDistributedBackend", "LocalBackend", "NVLinkBackend",
           "create_local_ddp", "create_distributed_ddp"]
