"""``GPT2Trainer`` — causal-LM fine-tuning under 3-D parallelism
(reference ``GPT2_Trainer.py:56-554``).

Keeps: constructor ``(model, train_loader, val_loader, config, pg_manager,
model_config)``, AdamW(wd 0.01) + CE(ignore −100), per-epoch validation,
``Train Loss/PPL`` prints, per-rank shard checkpoints
``{output_dir}/{best_model|final_model}_pp{p}_tp{t}.pt`` with the 4 top-level
keys.  Adds: real resume (``resume_from``), ZeRO-1, bf16, fused
lm_head+CE, shifted labels, correct tied-weight sync, no hard-coded sizes.
"""
from __future__ import annotations

import math
import os
import time

import torch
import torch.nn as nn

from .parallel.pipeline_parallel import PipelineDataLoader, PipelineTrainer
from .parallel.pipeline_parallel.schedule import unwrap_stage
from .utils.logger import StepLogger
from .trainer import _TrainerBase, build_optimizer
from .utils import checkpoint as ckpt


class GPT2Trainer(_TrainerBase):
    task_type = "clm"

    def __init__(self, model: nn.Module, train_loader, val_loader, config: dict, pg_manager,
                 model_config=None, tokenizer=None):
        super().__init__(model, train_loader, val_loader, config, pg_manager)
        self.model_config = model_config
        self.tokenizer = tokenizer
        self.vocab_size = model_config.vocab_size if model_config is not None \
            else int(config.get("vocab_size", 50257))
        self.optimizer = build_optimizer(model, config, pg_manager, default="adamw", weight_decay=0.01)
        self.criterion = nn.CrossEntropyLoss(ignore_index=-100)
        self.gpt2_stage = unwrap_stage(model)
        self.step_logger = None
        self.shift_labels = bool(config.get("shift_labels", True))
        if hasattr(self.gpt2_stage, "shift_labels"):
            self.gpt2_stage.shift_labels = self.shift_labels
        self._make_lr_sched(0)
        self.max_seq_length = int(config.get("max_seq_length", 512))
        if hasattr(self.gpt2_stage, "seq_len"):
            self.gpt2_stage.seq_len = self.max_seq_length
        self.pipeline_trainer = PipelineTrainer(
            model=self.model, device_mesh=pg_manager.device_mesh, pp_rank=self.pp_rank,
            pp_group=pg_manager.get_group("pp"), criterion=self.criterion, device=self.device,
            optimizer=self.optimizer, max_grad_norm=self.max_grad_norm,
            schedule_type=self._schedule_name(config), task_type="clm", vocab_size=self.vocab_size,
            loss_scale=config.get("loss_scale", "mean"), shift_labels=self.shift_labels,
            pg_manager=pg_manager, compute_dtype=self.dtype)
        self.micro_loader = PipelineDataLoader(self.train_loader, self.grad_acc_steps, task_type="clm")
        mb = self.micro_batch_size()
        self.tensor_shapes = self._shapes(mb)
        if config.get("resume_from"):
            self.resume(config["resume_from"])

    @staticmethod
    def _schedule_name(config: dict) -> str:
        """``virtual_pipeline_stages > 1`` implies the interleaved schedule."""
        name = config.get("schedule", "1f1b")
        if int(config.get("virtual_pipeline_stages", 1) or 1) > 1 and name == "1f1b":
            return "1f1b_interleaved"
        return name

    def _shapes(self, batch: int):
        if hasattr(self.gpt2_stage, "get_tensor_shapes"):
            self.gpt2_stage.seq_len = self.max_seq_length
            return tuple(self.gpt2_stage.get_tensor_shapes(batch))
        E = self.model_config.n_embd if self.model_config else 768
        return (batch, self.max_seq_length, E)

    # ------------------------------------------------------------------- fit
    def train_step(self):
        """One optimizer step (``grad_acc_steps`` micro-batches).  Works for any
        pp size: with pp=1 the 1F1B schedule degenerates to plain accumulation."""
        out = self.pipeline_trainer.train_step(self.micro_loader, self.tensor_shapes, self.device, self.dtype)
        if self.lr_sched:
            self.lr_sched.step()
        self.global_step += 1
        return out

    def fit(self):
        best_val_ppl = float("inf")
        for epoch in range(self.start_epoch, int(self.config["num_epochs"])):
            if self.global_rank == 0:
                print(f"\nEpoch {epoch + 1}/{self.config['num_epochs']}\n" + "-" * 50)
            t0 = time.time()
            train_loss, train_ppl = self._train_epoch(epoch)
            val_loss, val_ppl = self._validate_epoch(epoch)
            train_loss, train_ppl, val_loss, val_ppl = self._reduce_metrics_max(
                [train_loss, _finite(train_ppl), val_loss, _finite(val_ppl)])
            if self.global_rank == 0:
                print(f"Epoch {epoch + 1} Results:")
                print(f"  Train Loss: {train_loss:.4f} | Train PPL: {train_ppl:.2f}")
                print(f"  Val Loss:   {val_loss:.4f} | Val PPL:   {val_ppl:.2f}")
            self.logger.log(epoch=epoch + 1, train_loss=train_loss, train_ppl=train_ppl,
                            val_loss=val_loss, val_ppl=val_ppl, epoch_s=time.time() - t0)
            self.epoch_done = epoch + 1
            if val_ppl < best_val_ppl:
                best_val_ppl = val_ppl
                self._barrier()
                self._save_checkpoint("best_model.pt")
                self._barrier()
                if self.global_rank == 0:
                    print(f"  New best model saved (PPL: {best_val_ppl:.2f})")
        self._barrier()
        self._save_checkpoint("final_model.pt")
        self._barrier()
        if self.global_rank == 0:
            print("\nTraining complete.")
        if self.config.get("compute_generation_metrics", False):
            self._evaluate_generation_metrics()

    def _train_epoch(self, epoch: int):
        self.model.train()
        sampler = getattr(self.train_loader, "sampler", None)
        if hasattr(sampler, "set_epoch"):
            sampler.set_epoch(epoch)
        n_steps = max(len(self.train_loader) // self.grad_acc_steps, 1)
        tot, n = 0.0, 0
        log_every = int(self.config.get("log_interval", 0) or 0)
        # per-step records come from the rank that sees the loss (last stage, dp 0, tp 0); its host read of the
        # loss closes the step, so wall time around train_step is the step time
        reporter = self.is_last_stage and self.pg_manager.axis_rank("dp") == 0 and self.pg_manager.axis_rank("tp") == 0
        if reporter and self.step_logger is None and self.config.get("log_jsonl"):
            self.step_logger = StepLogger(self.config["log_jsonl"], rank=0)
        tokens_per_step = int(self.config["batch_size"]) * self.max_seq_length
        mc = self.model_config
        # model FLOPs per token (fwd + bwd): 6 x (12 E^2 per block + tied lm_head V E) + causal attention 6 T E per block
        flops_per_token = None if mc is None else (
            6.0 * (mc.n_layer * 12 * mc.n_embd ** 2 + mc.vocab_size * mc.n_embd)
            + mc.n_layer * 6.0 * self.max_seq_length * mc.n_embd)
        for step in range(n_steps):
            t_step = time.perf_counter()
            loss, ppl = self.train_step()
            if loss is not None:
                tot, n = tot + loss, n + 1
                if log_every and (step + 1) % log_every == 0 and reporter:
                    print(f"  step {step + 1}/{n_steps} loss {loss:.4f} ppl {ppl:.2f}", flush=True)
                if reporter and self.step_logger is not None:
                    dt = time.perf_counter() - t_step
                    self.step_logger.log(step=self.global_step, loss=loss, ppl=ppl, step_ms=dt * 1e3,
                                         tokens_per_s=tokens_per_step / dt,
                                         model_tflops=(flops_per_token * tokens_per_step / dt / 1e12
                                                       if flops_per_token else None),
                                         max_mem_gb=(torch.cuda.max_memory_allocated(self.device) / 2 ** 30
                                                     if str(self.device).startswith("cuda") else None),
                                         lr=self.optimizer.param_groups[0]["lr"] if hasattr(self.optimizer, "param_groups") else None)
            si = int(self.config.get("save_interval", 0) or 0)
            if si and self.global_step % si == 0:
                self._save_checkpoint(f"step{self.global_step}.pt")
            ei = int(self.config.get("eval_interval", 0) or 0)       # mid-epoch validation every `eval_interval` steps
            if ei and self.val_loader is not None and self.global_step % ei == 0:
                v_loss, v_ppl = self._validate_epoch(epoch)
                self.model.train()
                if v_loss is not None and reporter:
                    print(f"  step {self.global_step} Val Loss: {v_loss:.4f} | Val PPL: {_finite(v_ppl):.2f}", flush=True)
                    if self.step_logger is not None:
                        self.step_logger.log(step=self.global_step, val_loss=v_loss, val_ppl=_finite(v_ppl))
        if n == 0:
            return None, None
        avg = tot / n
        return avg, (math.exp(avg) if avg < 20 else float("inf"))

    @torch.no_grad()
    def _validate_epoch(self, epoch: int):
        if self.val_loader is None:
            return 0.0, 0.0
        bsz = getattr(self.val_loader, "batch_size", None) or self.micro_batch_size()
        loss, ppl = self.pipeline_trainer.evaluate(self.val_loader, self._shapes(bsz), self.device, self.dtype)
        return loss, ppl

    # ------------------------------------------------------------ checkpoints
    def _shard_name(self, filename: str) -> str:
        base, ext = os.path.splitext(filename)
        return f"{base}_pp{self.pp_rank}_tp{self.pg_manager.axis_rank('tp')}{ext}"

    def _save_checkpoint(self, filename: str):
        out_dir = self.config.get("output_dir", "./checkpoints")
        ckpt.save_shard(
            out_dir, filename, self.gpt2_stage, self.optimizer, self.config, self.pg_manager,
            extra={"epoch": getattr(self, "epoch_done", 0), "global_step": self.global_step,
                   "lr_sched": self.lr_sched.state_dict() if self.lr_sched else None,
                   "model_config": self.model_config.to_dict() if self.model_config is not None else None})

    def resume(self, prefix_path: str):
        """``prefix_path`` = ``{output_dir}/{name}.pt`` (without the ``_pp_tp`` suffix)."""
        meta = ckpt.load_shard(prefix_path, self.gpt2_stage, self.optimizer, self.pg_manager,
                               map_location=self.device)
        self.start_epoch = int(meta.get("epoch", 0))
        self.global_step = int(meta.get("global_step", 0))
        if self.lr_sched and meta.get("lr_sched"):
            self.lr_sched.load_state_dict(meta["lr_sched"])
        return meta

    # ------------------------------------------------------------- generation
    def _evaluate_generation_metrics(self):
        if self.is_pipeline or self.pg_manager.axis_size("tp") > 1 or self.tokenizer is None:
            if self.global_rank == 0:
                print("Generation metrics need a single-stage model and a tokenizer; skipped.")
            return None
        from .utils.metrics import evaluate_generation
        res = evaluate_generation(self.gpt2_stage, self.tokenizer, self.val_loader.dataset, self.device,
                                  num_samples=int(self.config.get("num_eval_samples", 50)),
                                  max_new_tokens=int(self.config.get("max_target_length", 128)))
        if self.global_rank == 0:
            print("Generation metrics:", res)
        return res


def _finite(x):
    if x is None:
        return None
    return x if math.isfinite(x) else 1e30
