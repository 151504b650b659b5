"""Per-rank sharded loading of GPT-2 weights.

``load_gpt2_distributed`` has the reference's signature and key schema
(``core/distributed_loading.py:203-376``): it opens a HuggingFace GPT-2
``.safetensors`` memory-mapped and materialises ONLY this rank's tensors —
``wte/wpe`` on the first stage, the owned layers' LayerNorms (replicated) and
TP slices of ``c_attn`` (column), ``attn.c_proj`` (row, bias on tp0), ``c_fc``
(column), ``mlp.c_proj`` (row, bias on tp0), and ``ln_f`` + a copy of ``wte``
on the last stage for the tied lm_head.  Linear weights are returned in
``nn.Linear`` layout ``[out, in]`` (HF stores Conv1D ``[in, out]``).

Difference: the fused QKV projection is sliced **by heads** — rank ``r`` gets
``[Q_r | K_r | V_r]`` — so the sharded model computes GPT-2.

``random_gpt2_state_dict`` produces the same dict from a seeded generator (no
checkpoint needed: benchmarks and tests use random weights), identical across
ranks for replicated tensors and consistent across TP slices.
"""
from __future__ import annotations

import zlib
from typing import Dict, Optional

import torch

from ..parallel.pipeline_parallel.utils import distribute_layers


def _coords(pg_manager):
    if pg_manager is None:
        return 0, 1, 0, 1
    return (pg_manager.axis_rank("tp"), pg_manager.axis_size("tp"),
            pg_manager.axis_rank("pp"), pg_manager.axis_size("pp"))


def _cfg(model_config):
    from ..models.gpt2.gpt2_config import GPT2Config
    if model_config is None:
        return GPT2Config()
    if isinstance(model_config, dict):
        return GPT2Config.from_dict(model_config)
    return model_config


def qkv_columns(E: int, tp_rank: int, tp_size: int) -> torch.Tensor:
    """Column indices of HF's fused [Q|K|V] owned by ``tp_rank`` (head-aware)."""
    w = E // tp_size
    base = torch.arange(tp_rank * w, (tp_rank + 1) * w)
    return torch.cat([base, base + E, base + 2 * E])


def shard_hf_tensors(get, config, tp_rank: int, tp_size: int, pp_rank: int, pp_size: int,
                     device, dtype=None) -> Dict[str, torch.Tensor]:
    """``get(name) -> Tensor`` fetches a full HF tensor (``wte.weight``,
    ``h.0.attn.c_attn.weight`` …, without the ``transformer.`` prefix)."""
    E = config.n_embd
    inner = config.n_inner
    is_first, is_last = pp_rank == 0, pp_rank == pp_size - 1
    layers = distribute_layers(config.n_layer, pp_size, getattr(config, 'pp_layer_counts', None))[pp_rank]
    out: Dict[str, torch.Tensor] = {}

    def put(name, t):
        t = t.to(device=device)
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        out[name] = t.contiguous()

    if is_first:
        put("wte.weight", get("wte.weight"))
        if not getattr(config, "use_rope", False):
            put("wpe.weight", get("wpe.weight"))
    cols = qkv_columns(E, tp_rank, tp_size)
    e0, e1 = tp_rank * (E // tp_size), (tp_rank + 1) * (E // tp_size)
    i0, i1 = tp_rank * (inner // tp_size), (tp_rank + 1) * (inner // tp_size)
    for i in layers:
        p = f"h.{i}"
        for ln in ("ln_1", "ln_2"):
            put(f"{p}.{ln}.weight", get(f"{p}.{ln}.weight"))
            put(f"{p}.{ln}.bias", get(f"{p}.{ln}.bias"))
        w = get(f"{p}.attn.c_attn.weight")                 # Conv1D [E, 3E]
        put(f"{p}.attn.c_attn.weight", w[:, cols].t())     # -> [3E/tp, E]
        put(f"{p}.attn.c_attn.bias", get(f"{p}.attn.c_attn.bias")[cols])
        w = get(f"{p}.attn.c_proj.weight")                 # [E, E]
        put(f"{p}.attn.c_proj.weight", w[e0:e1, :].t())    # -> [E, E/tp]
        if tp_rank == 0:
            put(f"{p}.attn.c_proj.bias", get(f"{p}.attn.c_proj.bias"))
        w = get(f"{p}.mlp.c_fc.weight")                    # [E, 4E]
        put(f"{p}.mlp.c_fc.weight", w[:, i0:i1].t())       # -> [4E/tp, E]
        put(f"{p}.mlp.c_fc.bias", get(f"{p}.mlp.c_fc.bias")[i0:i1])
        w = get(f"{p}.mlp.c_proj.weight")                  # [4E, E]
        put(f"{p}.mlp.c_proj.weight", w[i0:i1, :].t())     # -> [E, 4E/tp]
        if tp_rank == 0:
            put(f"{p}.mlp.c_proj.bias", get(f"{p}.mlp.c_proj.bias"))
    if is_last:
        put("ln_f.weight", get("ln_f.weight"))
        put("ln_f.bias", get("ln_f.bias"))
        if "wte.weight" not in out:
            put("wte.weight", get("wte.weight"))           # tied lm_head copy
    return out


def load_gpt2_distributed(checkpoint_path: str, pg_manager, config: Optional[dict] = None, device=None,
                          model_config=None, dtype=None, pp_rank: Optional[int] = None,
                          pp_size: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """``pp_rank`` / ``pp_size`` override the mesh coordinates (virtual pipeline stages load stage
    ``v * pp + rank`` of ``pp * V``)."""
    from safetensors import safe_open
    cfg = _cfg(model_config if model_config is not None else (config or {}).get("model_config"))
    tp_rank, tp_size, mesh_pp_rank, mesh_pp_size = _coords(pg_manager)
    pp_rank = mesh_pp_rank if pp_rank is None else pp_rank
    pp_size = mesh_pp_size if pp_size is None else pp_size
    device = device if device is not None else torch.device("cpu")
    with safe_open(checkpoint_path, framework="pt", device="cpu") as f:
        keys = set(f.keys())
        prefix = "transformer." if any(k.startswith("transformer.") for k in keys) else ""

        def get(name):
            return f.get_tensor(prefix + name)

        return shard_hf_tensors(get, cfg, tp_rank, tp_size, pp_rank, pp_size, device, dtype)


# --------------------------------------------------------------------------
# random init of the same shapes
# --------------------------------------------------------------------------
def hf_shapes(config) -> Dict[str, tuple]:
    E, I, V, P = config.n_embd, config.n_inner, config.vocab_size, config.n_positions
    shapes = {"wte.weight": (V, E), "wpe.weight": (P, E), "ln_f.weight": (E,), "ln_f.bias": (E,)}
    for i in range(config.n_layer):
        p = f"h.{i}"
        shapes.update({
            f"{p}.ln_1.weight": (E,), f"{p}.ln_1.bias": (E,), f"{p}.ln_2.weight": (E,), f"{p}.ln_2.bias": (E,),
            f"{p}.attn.c_attn.weight": (E, 3 * E), f"{p}.attn.c_attn.bias": (3 * E,),
            f"{p}.attn.c_proj.weight": (E, E), f"{p}.attn.c_proj.bias": (E,),
            f"{p}.mlp.c_fc.weight": (E, I), f"{p}.mlp.c_fc.bias": (I,),
            f"{p}.mlp.c_proj.weight": (I, E), f"{p}.mlp.c_proj.bias": (E,),
        })
    return shapes


def random_hf_tensor(name: str, shape, config, seed: int, device) -> torch.Tensor:
    """Deterministic GPT-2-style init of one full HF tensor (same on every rank)."""
    if name.endswith("ln_1.weight") or name.endswith("ln_2.weight") or name == "ln_f.weight":
        return torch.ones(shape, device=device)
    if name.endswith(".bias"):
        return torch.zeros(shape, device=device)
    g = torch.Generator(device=device)
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
    std = config.initializer_range
    if name.endswith("c_proj.weight"):
        std = std / (2 * config.n_layer) ** 0.5
    return torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std


def random_gpt2_state_dict(model_config, pg_manager=None, device=None, seed: int = 0, dtype=None,
                           tp_rank: Optional[int] = None, tp_size: Optional[int] = None,
                           pp_rank: Optional[int] = None, pp_size: Optional[int] = None):
    cfg = _cfg(model_config)
    c = _coords(pg_manager)
    tp_rank = c[0] if tp_rank is None else tp_rank
    tp_size = c[1] if tp_size is None else tp_size
    pp_rank = c[2] if pp_rank is None else pp_rank
    pp_size = c[3] if pp_size is None else pp_size
    device = device if device is not None else torch.device("cpu")
    shapes = hf_shapes(cfg)

    def get(name):
        return random_hf_tensor(name, shapes[name], cfg, seed, device)

    return shard_hf_tensors(get, cfg, tp_rank, tp_size, pp_rank, pp_size, device, dtype)


def save_random_hf_checkpoint(path: str, model_config, seed: int = 0, prefix: str = "") -> str:
    """Write a random-init HF-layout GPT-2 ``.safetensors`` (benchmarks/tests:
    there is no network to download the real one)."""
    from safetensors.torch import save_file
    cfg = _cfg(model_config)
    tensors = {prefix + n: random_hf_tensor(n, s, cfg, seed, torch.device("cpu")).contiguous()
               for n, s in hf_shapes(cfg).items()}
    save_file(tensors, path)
    return path
