"""``Trainer`` (ViT / classification) — reference ``trainer.py:57-362``.

Same constructor, ``fit()`` loop, epoch print format (the reference's launchers
regex-parse ``"Train Loss: x | Train Acc: y%"``) and end-of-fit rank-0
``model_checkpoint.pt``.  Differences: works for any mesh (missing axes = 1),
gradients are really synchronised over DP, metrics are accumulated on device,
``patience`` early stopping and ``warmup_steps`` are honoured, bf16 compute via
``config['dtype']``, optional ZeRO-1 (``config['zero_stage']=1``).
"""
from __future__ import annotations

import os
import time

import torch
import torch.distributed as dist
import torch.nn as nn

from .coordinators import resolve_device, resolve_dtype
from .optim import DistributedAdamW, FusedAdamW, WarmupLR, clip_grad_norm_global, zero_grads
from .parallel.data_parallel import DataParallel
from .parallel.pipeline_parallel import PipelineDataLoader, PipelineTrainer
from .parallel.pipeline_parallel.schedule import unwrap_stage
from .utils.logger import StepLogger


def build_optimizer(model, config: dict, pg_manager, default: str = "adam", weight_decay: float = 0.0):
    """Optimizer factory shared by both trainers.

    ``zero_stage>=1`` → ZeRO-1 :class:`DistributedAdamW`; CUDA → :class:`FusedAdamW`
    (native kernel); otherwise the stock torch optimizer the reference uses."""
    lr = float(config["learning_rate"])
    wd = float(config.get("weight_decay", weight_decay))
    kind = str(config.get("optimizer", default)).lower()
    if kind == "adam":
        wd = 0.0
    if int(config.get("zero_stage", 0)) >= 1:
        if not isinstance(model, DataParallel) or model.reduce_mode != "defer":
            raise ValueError("zero_stage=1 needs the model wrapped by the dp/3d strategy with zero_stage=1")
        return DistributedAdamW(model, lr=lr, weight_decay=wd, pg_manager=pg_manager,
                                max_grad_norm=float(config.get("max_grad_norm", 0.0)))
    p = next(model.parameters())
    if p.is_cuda or config.get("fused_optimizer", False):
        if isinstance(model, DataParallel) and all(b.flat_param is not None for b in model.buckets.values()):
            return FusedAdamW.from_data_parallel(model, lr=lr, weight_decay=wd)
        if not isinstance(model, DataParallel):
            from .coordinators.main_coordinator import resolve_main_grad_dtype
            return FusedAdamW(model.parameters(), lr=lr, weight_decay=wd, grad_dtype=resolve_main_grad_dtype(config))
    if kind == "adam":
        return torch.optim.Adam(model.parameters(), lr=lr)
    return torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=wd)


class _TrainerBase:
    task_type = "classification"

    def __init__(self, model: nn.Module, train_loader, val_loader, config: dict, pg_manager):
        self.model, self.train_loader, self.val_loader = model, train_loader, val_loader
        self.config, self.pg_manager = config, pg_manager
        if os.environ.get("QUINTNET_OUTPUT_DIR"):       # job launchers (tools/cluster.py) redirect the checkpoints
            config["output_dir"] = os.environ["QUINTNET_OUTPUT_DIR"]
        self.device = resolve_device(pg_manager, config)
        self.dtype = resolve_dtype(config)
        self.global_rank = dist.get_rank() if dist.is_initialized() else 0
        self.pp_size = pg_manager.axis_size("pp")
        self.pp_rank = pg_manager.axis_rank("pp")
        self.dp_size = pg_manager.axis_size("dp")
        self.is_pipeline = self.pp_size > 1
        self.is_first_stage = self.pp_rank == 0
        self.is_last_stage = self.pp_rank == self.pp_size - 1
        self.grad_acc_steps = int(config.get("grad_acc_steps", 1))
        self.max_grad_norm = float(config.get("max_grad_norm", 0.0))
        self.logger = StepLogger(config.get("log_jsonl"), self.global_rank)
        from . import ops
        from .coordinators.main_coordinator import resolve_main_grad_dtype
        if config.get("fp8_recipe"):
            ops.set_fp8_recipe(str(config["fp8_recipe"]))
        if config.get("fuse_wgrad_accumulation") is not None:
            ops.set_fused_wgrad_accumulation(bool(config["fuse_wgrad_accumulation"]))
        elif resolve_main_grad_dtype(config) is not None and os.environ.get("QUINTNET_FUSE_WGRAD_ACCUM") is None:
            # fp32 main grads: let the wgrad GEMMs accumulate straight into them (no bf16 round trip per micro-batch)
            ops.set_fused_wgrad_accumulation(True)
        self.start_epoch = 0
        self.global_step = 0

    def micro_batch_size(self) -> int:
        mb = int(self.config["batch_size"]) // (self.grad_acc_steps * self.dp_size)
        return max(mb, 1)

    def _make_lr_sched(self, steps_per_epoch: int):
        ws = int(self.config.get("warmup_steps", 0))
        self.lr_sched = WarmupLR(self.optimizer, ws) if ws > 0 else None

    def _reduce_metrics_max(self, values):
        t = torch.tensor([0.0 if v is None else float(v) for v in values], device=self.device)
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def _barrier(self):
        if dist.is_initialized():
            dist.barrier()


class Trainer(_TrainerBase):
    task_type = "classification"

    def __init__(self, model: nn.Module, train_loader, val_loader, config: dict, pg_manager):
        super().__init__(model, train_loader, val_loader, config, pg_manager)
        self.optimizer = build_optimizer(model, config, pg_manager, default="adam")
        self.criterion = nn.CrossEntropyLoss()
        self._make_lr_sched(len(train_loader) if hasattr(train_loader, "__len__") else 0)
        if self.is_pipeline:
            self._setup_pipeline_training()
        elif self.config.get("cuda_graphs", False):
            self._enable_model_graphs(self.micro_batch_size())

    def _setup_pipeline_training(self):
        self.pipeline_trainer = PipelineTrainer(
            model=self.model, device_mesh=self.pg_manager.device_mesh, pp_rank=self.pp_rank,
            pp_group=self.pg_manager.get_group("pp"), criterion=self.criterion, device=self.device,
            optimizer=self.optimizer, max_grad_norm=self.max_grad_norm,
            schedule_type=self.config.get("schedule", "1f1b"), task_type="classification",
            loss_scale=self.config.get("loss_scale", "mean"), pg_manager=self.pg_manager,
            compute_dtype=self.dtype)
        self.train_loader = PipelineDataLoader(self.train_loader, self.grad_acc_steps)
        mb = self.micro_batch_size()
        self.config["micro_batch_size"] = mb
        self.tensor_shapes = tuple(self.model.get_tensor_shapes(mb))
        if self.config.get("cuda_graphs", False):
            self._enable_stage_graphs(mb)

    def _enable_model_graphs(self, mb: int):
        """``cuda_graphs: true`` without a pipeline: capture the whole model's micro-batch forward + backward."""
        if not (torch.cuda.is_available() and str(self.device).startswith("cuda")):
            return
        is_ddp = hasattr(self.model, "bucket_manager") and hasattr(self.model, "model")
        inner = self.model.model if is_ddp else self.model
        shape = (mb, int(self.config.get("in_channels", 1)), int(self.config.get("img_size", 28)),
                 int(self.config.get("img_size", 28)))
        try:
            if is_ddp:
                self.model.require_backward_grad_sync = False
            inner.train()
            sample = torch.randn(shape, device=self.device, dtype=self.dtype)
            graphed = torch.cuda.make_graphed_callables(inner, (sample,), num_warmup_iters=3, allow_unused_input=True)
            if is_ddp:
                self.model.model = graphed
            else:
                self.model = graphed
            self.graphed = True
        except Exception as e:  # noqa: BLE001 — capture is an optimisation; report and stay eager
            print(f"[quintnet_b200] CUDA-graph capture of the model failed, running eagerly: {type(e).__name__}: {str(e)[:300]}",
                  flush=True)
            self.graphed = False
        finally:
            if is_ddp:
                self.model.require_backward_grad_sync = True
        zero_grads(self.model, self.optimizer)

    def _enable_stage_graphs(self, mb: int):
        """``cuda_graphs: true`` — graph the per-micro-batch stage body (forward + backward): see
        ``PipelineParallelWrapper.enable_cuda_graphs``.  The gradient buckets are re-zeroed afterwards (capture ran
        warm-up backward passes through the DP hooks)."""
        from .parallel.pipeline_parallel.wrapper import PipelineParallelWrapper
        m = self.model.model if (hasattr(self.model, "bucket_manager") and hasattr(self.model, "model")) else self.model
        if not isinstance(m, PipelineParallelWrapper):
            return
        if m.is_first_stage:
            shape = (mb, int(self.config.get("in_channels", 1)), int(self.config.get("img_size", 28)),
                     int(self.config.get("img_size", 28)))
        else:
            shape = self.tensor_shapes
        is_ddp = hasattr(self.model, "require_backward_grad_sync")
        if is_ddp:
            self.model.require_backward_grad_sync = False
        self.graphed = m.enable_cuda_graphs(shape, self.dtype)
        if is_ddp:
            self.model.require_backward_grad_sync = True
        zero_grads(self.model, self.optimizer)

    # -------------------------------------------------------------------- fit
    def fit(self):
        best_val, bad_epochs = float("inf"), 0
        patience = int(self.config.get("patience", 0) or 0)
        for epoch in range(self.start_epoch, int(self.config["num_epochs"])):
            if self.global_rank == 0:
                print(f"\nEpoch {epoch + 1}/{self.config['num_epochs']}\n" + "-" * 30)
            t0 = time.time()
            train_loss, train_acc = self._train_epoch(epoch)
            val_loss, val_acc = self._validate_epoch(epoch)
            if self.is_pipeline:
                train_loss, train_acc, val_loss, val_acc = self._reduce_metrics_max(
                    [train_loss, train_acc, val_loss, val_acc])
            if self.global_rank == 0:
                print(f"Train Loss: {train_loss:.4f} | Train Acc: {train_acc:.2f}%")
                print(f"Val Loss:   {val_loss:.4f} | Val Acc:   {val_acc:.2f}%")
            self.logger.log(epoch=epoch + 1, train_loss=train_loss, train_acc=train_acc,
                            val_loss=val_loss, val_acc=val_acc, epoch_s=time.time() - t0)
            if val_loss < best_val - 1e-6:
                best_val, bad_epochs = val_loss, 0
            else:
                bad_epochs += 1
                if patience and bad_epochs >= patience:
                    if self.global_rank == 0:
                        print(f"Early stopping after {epoch + 1} epochs (patience {patience}).")
                    break
        if self.global_rank == 0:
            print("\nTraining complete.")
            self._save_checkpoint(self.config.get("checkpoint_name", "model_checkpoint.pt"))
        self._barrier()

    # ------------------------------------------------------------------ train
    def _train_epoch(self, epoch: int):
        self.model.train()
        sampler = getattr(getattr(self.train_loader, "dataloader", self.train_loader), "sampler", None)
        if hasattr(sampler, "set_epoch"):
            sampler.set_epoch(epoch)
        if self.is_pipeline:
            n_steps = max(len(self.train_loader.dataloader) // self.grad_acc_steps, 1)
            tot_loss, tot_acc, n = 0.0, 0.0, 0
            for _ in range(n_steps):
                loss, acc = self.pipeline_trainer.train_step(self.train_loader, self.tensor_shapes,
                                                             self.device, self.dtype)
                if self.lr_sched:
                    self.lr_sched.step()
                self.global_step += 1
                if loss is not None:
                    tot_loss, tot_acc, n = tot_loss + loss, tot_acc + acc, n + 1
            return (tot_loss / n, tot_acc / n) if n else (None, None)
        return self._train_epoch_standard(epoch)

    def _train_epoch_standard(self, epoch: int):
        loss_acc = torch.zeros((), device=self.device)
        correct = torch.zeros((), device=self.device)
        count = torch.zeros((), device=self.device)
        nb = 0
        M = self.grad_acc_steps
        is_ddp = hasattr(self.model, "require_backward_grad_sync")
        # only whole accumulation groups reach an optimizer step: a ragged tail of len(loader) % M micro-batches is
        # dropped (like drop_last) instead of leaking un-synchronised gradients into the next epoch's first step
        n_total = len(self.train_loader) if hasattr(self.train_loader, "__len__") else None
        limit = (n_total // M) * M if (n_total is not None and n_total >= M) else n_total
        zero_grads(self.model, self.optimizer)
        for i, batch in enumerate(self.train_loader):
            if limit is not None and i >= limit and limit >= M:
                break
            images, labels = _unpack_classification(batch)
            images = images.to(self.device, non_blocking=True).to(self.dtype)
            labels = labels.to(self.device, non_blocking=True)
            last = (i + 1) % M == 0 or (limit is not None and limit < M and i + 1 == limit)   # fewer than M batches: one step
            if is_ddp:
                self.model.require_backward_grad_sync = last
            out = self.model(images)
            loss = self.criterion(out.float(), labels)
            (loss / M).backward()
            loss_acc += loss.detach()
            correct += (out.argmax(1) == labels).sum()
            count += labels.shape[0]
            nb += 1
            if last:
                self._optimizer_step()
        l, c, n = torch.stack([loss_acc, correct, count]).tolist()
        return l / max(nb, 1), 100.0 * c / max(n, 1)

    def _optimizer_step(self):
        if self.max_grad_norm > 0 and not getattr(self.optimizer, "handles_grad_clipping", False):
            clip_grad_norm_global(self.model.parameters(), self.max_grad_norm, self.pg_manager)
        self.optimizer.step()
        zero_grads(self.model, self.optimizer)
        if self.lr_sched:
            self.lr_sched.step()
        self.global_step += 1

    # --------------------------------------------------------------- validate
    @torch.no_grad()
    def _validate_epoch(self, epoch: int):
        if self.val_loader is None:
            return 0.0, 0.0
        self.model.eval()
        if self.is_pipeline:
            shapes = self.tensor_shapes
            return self.pipeline_trainer.evaluate(self.val_loader, shapes, self.device, self.dtype)
        loss_sum = torch.zeros((), device=self.device)
        correct = torch.zeros((), device=self.device)
        count = torch.zeros((), device=self.device)
        for batch in self.val_loader:
            images, labels = _unpack_classification(batch)
            images = images.to(self.device).to(self.dtype)
            labels = labels.to(self.device)
            out = self.model(images)
            loss_sum += nn.functional.cross_entropy(out.float(), labels, reduction="sum")
            correct += (out.argmax(1) == labels).sum()
            count += labels.shape[0]
        stats = torch.stack([loss_sum, correct, count])
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(stats)
        l, c, n = stats.tolist()
        return l / max(n, 1), 100.0 * c / max(n, 1)

    # -------------------------------------------------------------- checkpoint
    def _save_checkpoint(self, filename: str):
        stage = unwrap_stage(self.model)
        sd = stage.state_dict()
        if self.is_pipeline:
            sd = {"pipeline_stage": sd}
        path = os.path.join(self.config.get("output_dir", "."), filename)
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save({"model_state_dict": sd, "config": self.config}, path)
        print(f"Checkpoint saved to {path}")


def _unpack_classification(batch):
    if isinstance(batch, dict):
        if "image" in batch:
            return batch["image"], batch["label"]
        return batch["images"], batch["labels"]
    return batch[0], batch[1]
