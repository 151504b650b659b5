"""No-op single-process backend (reference ``backends/local_backend.py:9-22``)."""
from .base import DistributedBackend, Handle


class LocalBackend(DistributedBackend):
    def is_initialized(self) -> bool:
        return True

    def get_world_size(self, group=None) -> int:
        return 1

    def broadcast_tensor(self, tensor, src, group=None) -> None:
        return None

    def all_reduce_tensor(self, tensor, op="sum", group=None) -> None:
        return None

    def reduce_scatter_async(self, out, inp, scale, group=None) -> Handle:
        out.copy_(inp[: out.numel()].view_as(out))
        if scale != 1.0:
            out.mul_(scale)
        return Handle()
