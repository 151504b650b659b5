"""GPT-2 configuration (reference ``utils/GPT2/gpt2_config.py:22-168``)."""
from __future__ import annotations

from dataclasses import dataclass, asdict, fields
from typing import Optional


@dataclass
class GPT2Config:
    vocab_size: int = 50257
    n_positions: int = 1024
    n_embd: int = 768
    n_layer: int = 12
    n_head: int = 12
    n_inner: Optional[int] = None
    activation_function: str = "gelu"
    resid_pdrop: float = 0.1
    embd_pdrop: float = 0.1
    attn_pdrop: float = 0.1
    layer_norm_epsilon: float = 1e-5
    initializer_range: float = 0.02
    # extensions (not in the reference)
    use_rope: bool = False            # RoPE instead of learned wpe (K7)
    rope_base: float = 10000.0
    sequence_parallel: bool = False   # token-sharded TP (fused AG->GEMM / GEMM->RS)
    pp_layer_counts: Optional[list] = None   # explicit layers per pipeline stage (default: even split)
    fp8_linear: bool = False          # forward projections on e4m3 operands with row-wise scales (bf16 backward)
    activation_checkpointing: bool = False   # recompute each block in backward instead of keeping its activations

    def __post_init__(self):
        if self.n_inner is None:
            self.n_inner = 4 * self.n_embd
        if self.n_embd % self.n_head:
            raise ValueError("n_embd must be divisible by n_head")

    @property
    def head_dim(self) -> int:
        return self.n_embd // self.n_head

    @classmethod
    def from_dict(cls, config_dict: dict) -> "GPT2Config":
        names = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in config_dict.items() if k in names})

    @classmethod
    def from_model_config(cls, model_config: dict) -> "GPT2Config":
        return cls.from_dict(dict(model_config))

    def to_dict(self) -> dict:
        return asdict(self)

    @classmethod
    def gpt2_base(cls, **kw) -> "GPT2Config":
        return cls(n_embd=768, n_layer=12, n_head=12, **kw)

    @classmethod
    def gpt2_medium(cls, **kw) -> "GPT2Config":
        return cls(n_embd=1024, n_layer=24, n_head=16, **kw)

    @classmethod
    def gpt2_large(cls, **kw) -> "GPT2Config":
        return cls(n_embd=1280, n_layer=36, n_head=20, **kw)

    @classmethod
    def gpt2_xl(cls, **kw) -> "GPT2Config":
        return cls(n_embd=1600, n_layer=48, n_head=25, **kw)

    @classmethod
    def tiny(cls, **kw) -> "GPT2Config":
        """Test-size model."""
        d = dict(vocab_size=128, n_positions=64, n_embd=32, n_layer=4, n_head=4,
                 resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
        d.update(kw)
        return cls(**d)
