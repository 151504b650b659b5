"""Merge ``{prefix}_pp{P}_tp{T}.pt`` shards into one HuggingFace-format GPT-2
checkpoint.

CLI parity: reference ``merge_checkpoints.py:247-258``
(``--input_dir --output --prefix``).  The reference's key patterns never match
its own shard keys; this tool is driven by the real schema
(§5.4): TP-merge (column layers ``cat(dim=0)``, row layers ``cat(dim=1)``,
replicated → rank-0 copy), PP-merge (``blocks.X`` → ``h.{X+offset}``), HF
conversion (strip ``.proj``, transpose to Conv1D ``[in,out]``,
``embedding.``→``transformer.``, ``lm_head_weight``→``lm_head.weight``).
QKV shards are head-aware (``[Q_r|K_r|V_r]`` per rank) and are re-interleaved
into HF's ``[Q|K|V]``; ``--legacy_qkv`` merges reference-style contiguous blocks.
"""
from __future__ import annotations

import argparse
import os
import re
from typing import Dict, Optional, Tuple

import torch

from ..parallel.pipeline_parallel.utils import distribute_layers

_SHARD_RE = re.compile(r"_pp(\d+)_tp(\d+)\.pt$")


def load_shards(input_dir: str, prefix: str = "final_model") -> Dict[Tuple[int, int], dict]:
    shards = {}
    for fn in sorted(os.listdir(input_dir)):
        if not fn.startswith(prefix + "_pp"):
            continue
        m = _SHARD_RE.search(fn)
        if not m:
            continue
        shards[(int(m.group(1)), int(m.group(2)))] = torch.load(os.path.join(input_dir, fn),
                                                                 map_location="cpu", weights_only=False)
    if not shards:
        raise FileNotFoundError(f"no shards named {prefix}_pp*_tp*.pt in {input_dir}")
    return shards


def merge_tp_shards(tp_shards: Dict[int, dict], pp_rank: Optional[int] = None, legacy_qkv: bool = False) -> dict:
    """Merge the TP ranks of ONE pipeline stage (reference signature ``(tp_shards, pp_rank)``,
    ``merge_checkpoints.py:59``).  ``tp_shards[t]`` is a stage state dict or a whole shard payload holding one under
    ``'model_state_dict'``; ``pp_rank`` is informational."""
    tp_states = {t: (s["model_state_dict"] if isinstance(s, dict) and "model_state_dict" in s else s)
                 for t, s in tp_shards.items()}
    tp = len(tp_states)
    s0 = tp_states[0]
    if tp == 1:
        return dict(s0)
    out = {}
    for key, v0 in s0.items():
        parts = [tp_states[r][key] for r in range(tp) if key in tp_states[r]]
        if re.search(r"\.(c_attn)\.proj\.(weight|bias)$", key):
            if legacy_qkv:
                out[key] = torch.cat(parts, dim=0)
            else:  # each part is [Q_r;K_r;V_r] -> [Q_0..Q_n; K_0..K_n; V_0..V_n]
                thirds = [p.chunk(3, dim=0) for p in parts]
                out[key] = torch.cat([t[j] for j in range(3) for t in thirds], dim=0)
        elif re.search(r"\.c_fc\.proj\.(weight|bias)$", key):
            out[key] = torch.cat(parts, dim=0)
        elif re.search(r"\.c_proj\.proj\.weight$", key):
            out[key] = torch.cat(parts, dim=1)
        else:  # replicated (LN, embeddings, lm_head) or tp0-only (c_proj.bias)
            out[key] = v0
    return out


def merge_pp_stages(pp_stages: Dict[int, dict], config=None, counts=None, n_layer: Optional[int] = None) -> dict:
    """Reference signature ``(pp_stages, config)`` (``merge_checkpoints.py:100``): ``config`` is the model config dict
    (``n_layer``); an int is taken as the layer count directly.  ``counts`` = explicit layers per stage."""
    stage_states = pp_stages
    if n_layer is None:
        n_layer = config if isinstance(config, int) else int((config or {}).get("n_layer", 0))
    if not n_layer:
        n_layer = sum(len({k.split(".")[1] for k in s if k.startswith("blocks.")}) for s in stage_states.values())
    pp = len(stage_states)
    layer_map = distribute_layers(n_layer, pp, counts)
    out = {}
    for p in range(pp):
        for key, v in stage_states[p].items():
            m = re.match(r"blocks\.(\d+)\.(.*)$", key)
            if m:
                out[f"blocks.{layer_map[p][int(m.group(1))]}.{m.group(2)}"] = v
            elif key == "lm_head_weight" and "lm_head_weight" in out:
                continue
            else:
                out[key] = v
    return out


def convert_to_hf_format(merged_state: dict) -> dict:
    merged = merged_state
    hf = {}
    for key, v in merged.items():
        if key == "lm_head_weight":
            hf["lm_head.weight"] = v
            continue
        k = key
        k = re.sub(r"^embedding\.", "transformer.", k)
        k = re.sub(r"^blocks\.", "transformer.h.", k)
        k = re.sub(r"^ln_f\.", "transformer.ln_f.", k)
        is_linear_w = bool(re.search(r"\.(c_attn|c_proj|c_fc)\.proj\.weight$", k))
        k = re.sub(r"\.(c_attn|c_proj|c_fc)\.proj\.", r".\1.", k)
        hf[k] = v.t().contiguous() if is_linear_w else v
    if "lm_head.weight" not in hf and "transformer.wte.weight" in hf:
        hf["lm_head.weight"] = hf["transformer.wte.weight"]
    return hf


def merge_checkpoints(input_dir: str, output_path: str, prefix: str = "final_model",
                      legacy_qkv: bool = False) -> dict:
    output = output_path
    shards = load_shards(input_dir, prefix)
    pps = sorted({p for p, _ in shards})
    tps = sorted({t for _, t in shards})
    any_shard = next(iter(shards.values()))
    config = any_shard.get("config", {})
    mc = (any_shard.get("extra") or {}).get("model_config") or config.get("model_config") or {}
    stage_states = {}
    for p in pps:
        stage_states[p] = merge_tp_shards({t: shards[(p, t)] for t in tps}, p, legacy_qkv)
    n_layer = int(mc.get("n_layer", 0)) or sum(
        len({k.split(".")[1] for k in s if k.startswith("blocks.")}) for s in stage_states.values())
    counts = (any_shard.get("parallelism_info") or {}).get("pp_layer_counts")
    merged = merge_pp_stages(stage_states, n_layer, counts)
    hf = convert_to_hf_format(merged)
    os.makedirs(os.path.dirname(os.path.abspath(output)), exist_ok=True)
    torch.save({"model_state_dict": hf, "config": mc}, output)
    return hf


def main(argv=None):
    ap = argparse.ArgumentParser(description="Merge QuintNet TP/PP shards into a HF GPT-2 checkpoint")
    ap.add_argument("--input_dir", required=True)
    ap.add_argument("--output", required=True)
    ap.add_argument("--prefix", default="final_model")
    ap.add_argument("--legacy_qkv", action="store_true",
                    help="shards use the reference's contiguous (non head-aware) QKV split")
    a = ap.parse_args(argv)
    hf = merge_checkpoints(a.input_dir, a.output, a.prefix, a.legacy_qkv)
    print(f"merged {len(hf)} tensors -> {a.output}")


if __name__ == "__main__":
    main()
