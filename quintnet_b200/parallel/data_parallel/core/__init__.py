from .config import ReductionStrategy, BucketConfig, DistributedConfig
from .ddp import DataParallel
__all__ = ["DataParallel", "ReductionStrategy", "BucketConfig", "DistributedConfig"]
