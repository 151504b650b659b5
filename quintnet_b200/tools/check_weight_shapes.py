"""Verify that the sharded GPT-2 loader produces, for every (tp, pp) rank, exactly the tensor shapes the tensor-parallel
layers expect — the job of the reference's root script ``test_weight_shapes.py`` (which hard-codes one checkpoint
path, tp = 2 and GPT-2 small).

    python -m quintnet_b200.tools.check_weight_shapes --checkpoint model.safetensors --tp 2 --pp 2
    python -m quintnet_b200.tools.check_weight_shapes --model gpt2_medium --tp 4          # seeded random weights
"""
from __future__ import annotations

import argparse

import torch

from ..core.distributed_loading import hf_shapes, random_hf_tensor, shard_hf_tensors
from ..models.gpt2 import GPT2Config, GPT2Stage
from ..parallel.pipeline_parallel.utils import distribute_layers


def expected_shapes(cfg: GPT2Config, tp: int, pp_rank: int, pp: int) -> dict:
    E, I = cfg.n_embd, cfg.n_inner
    layers = distribute_layers(cfg.n_layer, pp, getattr(cfg, "pp_layer_counts", None))[pp_rank]
    out = {}
    if pp_rank == 0:
        out.update({"wte.weight": (cfg.vocab_size, E), "wpe.weight": (cfg.n_positions, E)})
    for i in layers:
        p = f"h.{i}"
        out.update({f"{p}.ln_1.weight": (E,), f"{p}.ln_1.bias": (E,), f"{p}.ln_2.weight": (E,), f"{p}.ln_2.bias": (E,),
                    f"{p}.attn.c_attn.weight": (3 * E // tp, E), f"{p}.attn.c_attn.bias": (3 * E // tp,),
                    f"{p}.attn.c_proj.weight": (E, E // tp),
                    f"{p}.mlp.c_fc.weight": (I // tp, E), f"{p}.mlp.c_fc.bias": (I // tp,),
                    f"{p}.mlp.c_proj.weight": (E, I // tp)})
    if pp_rank == pp - 1:
        out.update({"ln_f.weight": (E,), "ln_f.bias": (E,)})
    return out


def check(cfg: GPT2Config, get, tp: int, pp: int, verbose: bool = True) -> int:
    """Returns the number of checked tensors; raises AssertionError on the first mismatch."""
    n = 0
    for pp_rank in range(pp):
        for tp_rank in range(tp):
            sd = shard_hf_tensors(get, cfg, tp_rank, tp, pp_rank, pp, torch.device("cpu"), None)
            for name, shape in expected_shapes(cfg, tp, pp_rank, pp).items():
                assert name in sd, f"pp{pp_rank} tp{tp_rank}: {name} missing from the shard"
                assert tuple(sd[name].shape) == tuple(shape), \
                    f"pp{pp_rank} tp{tp_rank}: {name} is {tuple(sd[name].shape)}, layers expect {tuple(shape)}"
                n += 1
            # and the stage really builds from it
            stage = GPT2Stage.from_sharded_state_dict(sd, cfg, pp_rank, pp, tp_rank, tp, None, None, torch.device("cpu"))
            if verbose:
                print(f"pp{pp_rank} tp{tp_rank}: {len(sd):4d} tensors, stage has "
                      f"{sum(p.numel() for p in stage.parameters()) / 1e6:8.2f} M parameters  OK")
    return n


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint", default=None, help="HF-layout GPT-2 .safetensors (random weights if omitted)")
    ap.add_argument("--model", default="gpt2_base", choices=["gpt2_base", "gpt2_medium", "gpt2_large", "gpt2_xl"])
    ap.add_argument("--tp", type=int, default=2)
    ap.add_argument("--pp", type=int, default=1)
    a = ap.parse_args(argv)
    cfg = getattr(GPT2Config, a.model)()
    if a.checkpoint:
        from safetensors import safe_open
        with safe_open(a.checkpoint, framework="pt", device="cpu") as f:
            keys = set(f.keys())
            prefix = "transformer." if any(k.startswith("transformer.") for k in keys) else ""
            n = check(cfg, lambda name: f.get_tensor(prefix + name), a.tp, a.pp)
    else:
        shapes = hf_shapes(cfg)
        n = check(cfg, lambda name: random_hf_tensor(name, shapes[name], cfg, 0, torch.device("cpu")), a.tp, a.pp)
    print(f"all {n} shard shapes match (tp={a.tp}, pp={a.pp}, {a.model})")


if __name__ == "__main__":
    main()
