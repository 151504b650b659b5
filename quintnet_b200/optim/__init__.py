from .fused_adamw import FusedAdamW
from .distributed_adamw import DistributedAdamW
from .zero import wrap_zero1, zero1_optimizer
from .utils import clip_grad_norm_global, global_grad_norm, zero_grads, WarmupLR

__all__ = ["FusedAdamW", "DistributedAdamW", "wrap_zero1", "zero1_optimizer",
           "clip_grad_norm_global", "global_grad_norm", "zero_grads", "WarmupLR"]
