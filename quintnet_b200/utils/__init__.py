from .checkpoint import save_checkpoint, load_checkpoint, save_shard, load_shard
from .logger import Logger, setup_rank_logging, StepLogger, setup_logger, log_rank_0
from .profiling import DeviceTimer, profile_time, profile_memory, nvtx_range
from .memory import get_memory_usage, clear_cache
from .data import (CustomDataset, SummarizationDataset, SummarizationCollator, SummarizationDataLoader,
                   SyntheticMNIST, SyntheticTokens, make_loader, mnist_transform, pil_to_tensor_transform)
from .metrics import compute_rouge_bleu, generate_summary, evaluate_generation
from .utils import setup_device, count_parameters
from ..models.vit import Model          # reference utils/__init__.py exports the ViT ``Model``
