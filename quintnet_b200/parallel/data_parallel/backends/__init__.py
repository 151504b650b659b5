from .base import DistributedBackend, Handle
from .torch_backend import TorchDistributedBackend
from .local_backend import LocalBackend
from .nccl import NVLinkBackend
__all__ = ["DistributedBackend", "Handle", "TorchDistributedBackend", "LocalBackend", "NVLinkBackend"]
