"""The seven concrete coordinators (reference ``coordinators/*.py``)."""
from __future__ import annotations


import torch.nn as nn

from .main_coordinator import (BaseCoordinator, compose, make_tied_group, resolve_device, resolve_dtype)


class _Composed(BaseCoordinator):
    use_tp = use_pp = use_dp = False

    def parallelize(self) -> nn.Module:
        return compose(self.model, self.pg_manager, self.config, use_tp=self.use_tp,
                       use_pp=self.use_pp, use_dp=self.use_dp)


class DataParallelCoordinator(_Composed):
    use_dp = True


class TensorParallelCoordinator(_Composed):
    use_tp = True


class PipelineParallelCoordinator(_Composed):
    use_pp = True


class DPTCoordinator(_Composed):
    use_tp = use_dp = True


class DPPCoordinator(_Composed):
    use_pp = use_dp = True


class TPPCoordinator(_Composed):
    use_tp = use_pp = True


class Hybrid3DCoordinator(BaseCoordinator):
    """3-D composition.  ``is_staged=True`` builds only this rank's GPT-2 stage
    from a sharded load (``checkpoint_path``) or a seeded random init
    (``checkpoint_path`` None / ``'random'``) — reference
    ``coordinators/hybrid_3d_coordinator.py:71-168``."""

    def parallelize(self) -> nn.Module:
        if getattr(self, "is_staged", False):
            return self._parallelize_staged()
        return compose(self.model, self.pg_manager, self.config, use_tp=True, use_pp=True, use_dp=True)

    def _parallelize_staged(self) -> nn.Module:
        from ..core.distributed_loading import load_gpt2_distributed, random_gpt2_state_dict
        from ..models.gpt2 import GPT2Config, GPT2Stage
        pg, cfg = self.pg_manager, self.config
        device, dtype = resolve_device(pg, cfg), resolve_dtype(cfg)
        mc = cfg.get("model_config", {})
        model_config = mc if isinstance(mc, GPT2Config) else GPT2Config.from_dict(dict(mc))
        if cfg.get("sequence_parallel") is not None:
            model_config.sequence_parallel = bool(cfg["sequence_parallel"])
        elif pg.axis_size("tp") > 1 and "batch_size" in cfg:
            # not specified: token-sharded TP (the fused AG->GEMM / GEMM->RS kernels need it) whenever the micro-batch
            # splits evenly over the tp ranks
            mb = int(cfg["batch_size"]) // max(int(cfg.get("grad_acc_steps", 1)) * pg.axis_size("dp"), 1)
            model_config.sequence_parallel = mb >= pg.axis_size("tp") and mb % pg.axis_size("tp") == 0
        pp = pg.axis_size("pp")
        V = int(cfg.get("virtual_pipeline_stages", 1) or 1) if pp > 1 else 1
        n_stages = pp * V                     # virtual stage v * pp + rank lives on this rank (interleaved 1F1B)
        bal = cfg.get("pp_balance")
        if bal is not None and pp > 1:
            from ..parallel.pipeline_parallel.utils import balanced_gpt2_split
            model_config.pp_layer_counts = (balanced_gpt2_split(model_config.n_layer, n_stages)
                                            if bal == "auto" else [int(c) for c in bal])
        ckpt = getattr(self, "checkpoint_path", None)
        tied = make_tied_group(pg)
        chunks = []
        for v in range(V):
            sid = v * pp + pg.axis_rank("pp")
            if ckpt in (None, "", "random"):
                sd = random_gpt2_state_dict(model_config, pg, device=device, seed=int(cfg.get("seed", 0)),
                                            dtype=dtype, pp_rank=sid, pp_size=n_stages)
            else:
                sd = load_gpt2_distributed(ckpt, pg, cfg, device, model_config, dtype=dtype, pp_rank=sid,
                                           pp_size=n_stages)
            stage = GPT2Stage.from_sharded_state_dict(
                sd, model_config, sid, n_stages, pg.axis_rank("tp"), pg.axis_size("tp"), pg.get_group("tp"),
                pg.get_group("pp"), device, dtype=dtype, tied_group=tied)
            stage.seq_len = int(cfg.get("max_seq_length", model_config.n_positions))
            del sd
            chunks.append(stage)
        self.model_config = model_config
        if V > 1:
            return compose(None, pg, cfg, use_tp=False, use_pp=True, use_dp=True, stage_chunks=chunks)
        return compose(None, pg, cfg, use_tp=False, use_pp=True, use_dp=True, stage_module=chunks[0])
