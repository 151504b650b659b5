"""ViT for MNIST-shaped inputs.

Same module tree / state-dict keys as the reference model
(``utils/model.py:45-398``): ``embedding.{patch_embed.proj,cls_token,pos_embed}``,
``blocks.{i}.{norm1,attention.{Q,K,V},norm2,mlp.{down_proj,up_proj}}``,
``classification_head.down_proj`` — so its checkpoints load here.  The math is
identical (pre-LN, attention without an output projection, ReLU MLP); the
implementation routes through :mod:`quintnet_b200.ops` so the sm_100a kernels
are used on B200.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class Attention(nn.Module):
    def __init__(self, hidden_dim: int, n_heads: int):
        super().__init__()
        if hidden_dim % n_heads:
            raise ValueError("hidden_dim must be divisible by n_heads")
        self.hidden_dim, self.n_heads = hidden_dim, n_heads
        self.head_dim = hidden_dim // n_heads
        self.Q = nn.Linear(hidden_dim, hidden_dim)
        self.K = nn.Linear(hidden_dim, hidden_dim)
        self.V = nn.Linear(hidden_dim, hidden_dim)
        self.tp_group = None   # set for the 'paired' TP rewrite (local heads, one gather)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, T, _ = x.shape
        q, k, v = self.Q(x), self.K(x), self.V(x)
        # heads are inferred from the feature width so a column-parallel rewrite
        # without gather (local heads only) also works
        h = q.shape[-1] // self.head_dim
        q = q.view(B, T, h, self.head_dim).transpose(1, 2)
        k = k.view(B, T, h, self.head_dim).transpose(1, 2)
        v = v.view(B, T, h, self.head_dim).transpose(1, 2)
        o = ops.attention(q, k, v, causal=False, dropout_p=0.0, training=self.training)
        o = o.transpose(1, 2).reshape(B, T, h * self.head_dim)
        if o.shape[-1] != self.hidden_dim and self.tp_group is not None:
            from ..core.communication import All_Gather
            o = All_Gather.apply(o, self.tp_group, "slice", -1)
        return o


class MLP(nn.Module):
    def __init__(self, input_dim: int):
        super().__init__()
        self.input_dim = input_dim
        self.down_proj = nn.Linear(input_dim, input_dim * 4)
        self.relu = nn.ReLU()
        self.up_proj = nn.Linear(input_dim * 4, input_dim)

    def forward(self, x):
        return self.up_proj(self.relu(self.down_proj(x)))


class PatchEmbedding(nn.Module):
    def __init__(self, img_size=32, patch_size=4, in_channels=3, embed_dim=64):
        super().__init__()
        if img_size % patch_size:
            raise ValueError("img_size must be divisible by patch_size")
        self.img_size, self.patch_size = img_size, patch_size
        side = img_size // patch_size
        self.n_patches = side * side
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class TransformerBlock(nn.Module):
    def __init__(self, hidden_dim: int, n_heads: int):
        super().__init__()
        self.hidden_dim, self.n_heads = hidden_dim, n_heads
        self.norm1 = nn.LayerNorm(hidden_dim)
        self.attention = Attention(hidden_dim, n_heads)
        self.norm2 = nn.LayerNorm(hidden_dim)
        self.mlp = MLP(hidden_dim)

    def forward(self, x):
        x = x + self.attention(ops.layer_norm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps))
        x = x + self.mlp(ops.layer_norm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps))
        return x


class ClassificationHead(nn.Module):
    def __init__(self, hidden_dim: int, num_classes: int = 10):
        super().__init__()
        self.down_proj = nn.Linear(hidden_dim, num_classes)

    def forward(self, x):
        return self.down_proj(x[:, 0])


class ViTEmbedding(nn.Module):
    def __init__(self, img_size=28, patch_size=4, in_channels=1, hidden_dim=64):
        super().__init__()
        self.img_size, self.patch_size = img_size, patch_size
        self.in_channels, self.hidden_dim = in_channels, hidden_dim
        self.patch_embed = PatchEmbedding(img_size, patch_size, in_channels, hidden_dim)
        n = self.patch_embed.n_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, hidden_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, hidden_dim))
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)

    def forward(self, x):
        x = self.patch_embed(x)
        cls = self.cls_token.expand(x.shape[0], -1, -1).to(x.dtype)
        return torch.cat([cls, x], dim=1) + self.pos_embed.to(x.dtype)


class Model(nn.Module):
    """``embedding`` / ``blocks`` / ``classification_head`` are the contract the
    pipeline wrapper relies on (reference ``utils/model.py:352-358``)."""

    def __init__(self, img_size=28, patch_size=4, hidden_dim=64, in_channels=1, n_heads=4,
                 depth=4, num_classes=10):
        super().__init__()
        self.img_size, self.patch_size, self.hidden_dim = img_size, patch_size, hidden_dim
        self.depth, self.n_heads, self.num_classes = depth, n_heads, num_classes
        self.embedding = ViTEmbedding(img_size, patch_size, in_channels, hidden_dim)
        self.blocks = nn.ModuleList([TransformerBlock(hidden_dim, n_heads) for _ in range(depth)])
        self.classification_head = ClassificationHead(hidden_dim, num_classes)

    def forward(self, x):
        x = self.embedding(x)
        for blk in self.blocks:
            x = blk(x)
        return self.classification_head(x)

    def get_tensor_shapes(self, batch_size: int):
        return (batch_size, self.embedding.pos_embed.shape[1], self.hidden_dim)


ViT = Model
