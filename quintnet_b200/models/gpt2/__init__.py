from .gpt2_config import GPT2Config
from .gpt2_embeddings import GPT2Embedding
from .gpt2_attention import GPT2Attention
from .gpt2_mlp import GPT2MLP
from .gpt2_block import GPT2Block
from .gpt2_stage import GPT2Stage
from .gpt2_model import GPT2Model

__all__ = ["GPT2Config", "GPT2Embedding", "GPT2Attention", "GPT2MLP", "GPT2Block", "GPT2Stage", "GPT2Model"]
