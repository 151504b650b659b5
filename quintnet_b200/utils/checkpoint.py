"""Checkpoint save / load / resume.

Shard files keep the reference's naming and top-level keys
(``GPT2_Trainer.py:453-507``): ``{dir}/{name}_pp{p}_tp{t}.pt`` holding
``model_state_dict``, ``optimizer_state_dict``, ``config``, ``parallelism_info``.
Only dp-rank 0 writes the model shard (the reference lets all DP replicas
overwrite the same file); under ZeRO-1 every dp rank additionally writes its
optimizer shard ``…_dp{d}.opt.pt``.  Resume (absent from the reference,
SURVEY §5.4) restores model, optimizer (incl. ZeRO shards), step/epoch and RNG.
"""
from __future__ import annotations

import os
import random
from typing import Optional

import numpy as np
import torch


def _names(out_dir: str, filename: str, pg_manager, stage_id: Optional[int] = None):
    """``stage_id`` overrides the pp index in the file name: with V virtual pipeline stages per rank every model
    chunk is written as its own shard ``…_pp{v * pp + rank}_tp{t}`` of a (pp·V)-stage pipeline, which is exactly
    the layout the merge tool already understands."""
    base, ext = os.path.splitext(filename)
    pp, tp, dp = pg_manager.axis_rank("pp"), pg_manager.axis_rank("tp"), pg_manager.axis_rank("dp")
    pp = pp if stage_id is None else stage_id
    shard = os.path.join(out_dir, f"{base}_pp{pp}_tp{tp}{ext or '.pt'}")
    opt = os.path.join(out_dir, f"{base}_pp{pp}_tp{tp}_dp{dp}.opt{ext or '.pt'}")
    return shard, opt


def _chunks_of(stage_module):
    """[(stage id, pipeline size, module)] — one entry, or V entries for a rank that owns V virtual stages."""
    return None if not hasattr(stage_module, "chunks") else list(stage_module.chunks)


def rng_state():
    st = {"python": random.getstate(), "numpy": np.random.get_state(), "torch": torch.get_rng_state()}
    if torch.cuda.is_available():
        st["cuda"] = torch.cuda.get_rng_state()
    return st


def set_rng_state(st):
    random.setstate(st["python"])
    np.random.set_state(st["numpy"])
    torch.set_rng_state(st["torch"])
    if "cuda" in st and torch.cuda.is_available():
        torch.cuda.set_rng_state(st["cuda"])


def save_shard(out_dir: str, filename: str, stage_module, optimizer, config, pg_manager,
               extra: Optional[dict] = None, _stage: Optional[tuple] = None) -> str:
    chunks = _chunks_of(stage_module)
    if chunks is not None:                      # virtual pipeline stages: one shard per chunk, optimizer with chunk 0
        pp, r, V = pg_manager.axis_size("pp"), pg_manager.axis_rank("pp"), len(chunks)
        first = None
        for v, chunk in enumerate(chunks):
            path = save_shard(out_dir, filename, chunk, optimizer if v == 0 else None, config, pg_manager, extra,
                              _stage=(v * pp + r, pp * V, V))
            first = first or path
        return first
    os.makedirs(out_dir, exist_ok=True)
    shard, opt_path = _names(out_dir, filename, pg_manager, None if _stage is None else _stage[0])
    sharded_opt = getattr(optimizer, "handles_grad_clipping", False) and hasattr(optimizer, "shards")
    info = {"pp_rank": pg_manager.axis_rank("pp") if _stage is None else _stage[0],
            "tp_rank": pg_manager.axis_rank("tp"),
            "pp_size": pg_manager.axis_size("pp") if _stage is None else _stage[1],
            "virtual_pipeline_stages": 1 if _stage is None else _stage[2],
            "tp_size": pg_manager.axis_size("tp"),
            "dp_size": pg_manager.axis_size("dp"), "zero_stage": 1 if sharded_opt else 0,
            "pp_layer_counts": getattr(getattr(stage_module, "config", None), "pp_layer_counts", None)}
    if pg_manager.axis_rank("dp") == 0:
        payload = {
            "model_state_dict": {k: v.detach().cpu() for k, v in stage_module.state_dict().items()},
            "optimizer_state_dict": None if (optimizer is None or sharded_opt) else optimizer.state_dict(),
            "config": _plain(config),
            "parallelism_info": info,
            "extra": extra or {},
            "rng_state": rng_state(),
        }
        torch.save(payload, shard)
    if sharded_opt:
        torch.save({"optimizer_state_dict": optimizer.state_dict(), "parallelism_info": info}, opt_path)
    return shard


def load_shard(prefix_path: str, stage_module, optimizer, pg_manager, map_location="cpu",
               strict: bool = True, _stage: Optional[tuple] = None) -> dict:
    chunks = _chunks_of(stage_module)
    if chunks is not None:
        pp, r, V = pg_manager.axis_size("pp"), pg_manager.axis_rank("pp"), len(chunks)
        meta = {}
        for v, chunk in enumerate(chunks):
            m = load_shard(prefix_path, chunk, optimizer if v == 0 else None, pg_manager, map_location, strict,
                           _stage=(v * pp + r, pp * V, V))
            meta = meta or m
        return meta
    out_dir, filename = os.path.split(prefix_path)
    shard, opt_path = _names(out_dir or ".", filename, pg_manager, None if _stage is None else _stage[0])
    payload = torch.load(shard, map_location="cpu", weights_only=False)
    info = payload["parallelism_info"]
    want_pp = pg_manager.axis_size("pp") if _stage is None else _stage[1]
    if info["pp_size"] != want_pp or info["tp_size"] != pg_manager.axis_size("tp"):
        raise ValueError(f"checkpoint was written for pp={info['pp_size']} tp={info['tp_size']}")
    stage_module.load_state_dict(payload["model_state_dict"], strict=strict)
    if optimizer is not None:
        if info.get("zero_stage", 0) >= 1 and os.path.exists(opt_path):
            optimizer.load_state_dict(torch.load(opt_path, map_location="cpu", weights_only=False)
                                      ["optimizer_state_dict"])
        elif payload.get("optimizer_state_dict") is not None:
            optimizer.load_state_dict(payload["optimizer_state_dict"])
        _refresh_masters(optimizer)
    if payload.get("rng_state") is not None:
        try:
            set_rng_state(payload["rng_state"])
        except Exception:
            pass
    return payload.get("extra", {})


def _refresh_masters(optimizer) -> None:
    """fp32 master copies are part of the optimizer state; nothing to do unless
    the optimizer was freshly built from loaded low-precision parameters."""
    return None


def _plain(obj):
    if isinstance(obj, dict):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    if isinstance(obj, torch.dtype):
        return str(obj).replace("torch.", "")
    if hasattr(obj, "to_dict"):
        return obj.to_dict()
    return obj


# ---- single-process helpers with the signatures of the reference's ``utils/checkpoint.py:10-38`` (TODO stubs there) ----
def save_checkpoint(model, optimizer, epoch: int, checkpoint_dir, is_best: bool = False, **kwargs) -> str:
    """Write ``{checkpoint_dir}/checkpoint_epoch_{epoch}.pt`` (``epoch``, ``model_state_dict``,
    ``optimizer_state_dict`` + any extra ``kwargs`` such as ``loss=``) and, with ``is_best``, a copy as
    ``best_model.pt``.  Returns the path.  (The ``(model, optimizer, epoch, loss, path)`` variant of the reference's
    ``utils/utils.py`` lives in :mod:`quintnet_b200.utils.utils`.)"""
    checkpoint_dir = str(checkpoint_dir)
    os.makedirs(checkpoint_dir, exist_ok=True)
    payload = {"epoch": epoch, "model_state_dict": model.state_dict(),
               "optimizer_state_dict": None if optimizer is None else optimizer.state_dict(), **kwargs}
    path = os.path.join(checkpoint_dir, f"checkpoint_epoch_{epoch}.pt")
    torch.save(payload, path)
    if is_best:
        torch.save(payload, os.path.join(checkpoint_dir, "best_model.pt"))
    return path


def load_checkpoint(checkpoint_path, model, optimizer=None, strict: bool = True, map_location="cpu") -> dict:
    """Load a file written by :func:`save_checkpoint` (or any ``{'model_state_dict': …}`` file) into ``model`` /
    ``optimizer``; returns the remaining entries (``epoch``, ``loss``, …)."""
    ck = torch.load(str(checkpoint_path), map_location=map_location, weights_only=False)
    model.load_state_dict(ck["model_state_dict"], strict=strict)
    if optimizer is not None and ck.get("optimizer_state_dict") is not None:
        optimizer.load_state_dict(ck["optimizer_state_dict"])
    return {k: v for k, v in ck.items() if k not in ("model_state_dict", "optimizer_state_dict")}
