from .vit import Model, ViT
from .gpt2 import (GPT2Config, GPT2Embedding, GPT2Attention, GPT2MLP, GPT2Block, GPT2Stage, GPT2Model)

__all__ = ["Model", "ViT", "GPT2Config", "GPT2Embedding", "GPT2Attention", "GPT2MLP", "GPT2Block",
           "GPT2Stage", "GPT2Model"]
