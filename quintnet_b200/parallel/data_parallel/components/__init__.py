from .bucket import GradientBucket
from .bucket_manager import BucketManager
from .gradient_reducer import GradientReducer
from .parameter_broadcaster import ParameterBroadcaster
__all__ = ["GradientBucket", "BucketManager", "GradientReducer", "ParameterBroadcaster"]
