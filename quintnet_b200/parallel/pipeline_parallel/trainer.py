"""``PipelineTrainer`` — runs one optimizer step (or an evaluation pass) of a
pipeline stage under a chosen schedule.

Signature parity: reference ``parallelism/pipeline_parallel/trainer.py:31-281``
(``model, device_mesh, pp_rank, pp_group, criterion, device, optimizer,
max_grad_norm, schedule_type, task_type, vocab_size``); extra keyword arguments
select the corrected semantics documented in :mod:`.schedule`.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist

from ... import ops
from ...core.communication import drain_pipeline_sends, pipeline_communicate
from .schedule import AllFwdAllBwdSchedule, OneFOneBSchedule, unwrap_stage


class PipelineTrainer:
    def __init__(self, model, device_mesh, pp_rank: int, pp_group, criterion, device,
                 optimizer=None, max_grad_norm: float = 1.0, schedule_type: str = "1f1b",
                 task_type: str = "classification", vocab_size: Optional[int] = None,
                 loss_scale: str = "mean", shift_labels: bool = True, pg_manager=None,
                 compute_dtype: Optional[torch.dtype] = None):
        self.model, self.device_mesh = model, device_mesh
        self.rank, self.pp_group = pp_rank, pp_group
        self.world_size = dist.get_world_size(pp_group) if pp_group is not None else 1
        self.is_first_stage = pp_rank == 0
        self.is_last_stage = pp_rank == self.world_size - 1
        self.criterion, self.device, self.optimizer = criterion, device, optimizer
        self.max_grad_norm = float(max_grad_norm)
        self.task_type, self.vocab_size = task_type, vocab_size
        self.loss_scale, self.shift_labels = loss_scale, shift_labels
        self.pg_manager = pg_manager
        self.batch_labels, self.batch_predictions = [], []     # reference attributes (never filled there either)
        self.compute_dtype = compute_dtype
        stage = unwrap_stage(model)
        self.stage = stage
        self.fused_loss = bool(getattr(stage, "supports_fused_loss", False)) and task_type == "clm"
        if self.fused_loss and hasattr(stage, "shift_labels"):
            stage.shift_labels = shift_labels
        if schedule_type == "afab":
            self.schedule = AllFwdAllBwdSchedule(self)
        elif schedule_type == "1f1b":
            self.schedule = OneFOneBSchedule(self)
        elif schedule_type == "1f1b_interleaved":
            from .interleaved import InterleavedOneFOneBSchedule
            self.schedule = InterleavedOneFOneBSchedule(self)
        else:
            raise ValueError(f"unknown schedule_type {schedule_type!r} (use '1f1b', 'afab' or '1f1b_interleaved')")
        if hasattr(stage, "chunks") and schedule_type != "1f1b_interleaved":
            raise ValueError("virtual pipeline stages need schedule '1f1b_interleaved'")
        self.schedule_type = schedule_type
        self._warmup_p2p()

    def _warmup_p2p(self) -> None:
        """Exchange one tiny tensor with each pipeline neighbour so the (lazy) p2p communicators exist
        before the first timed / fused step — keeps ranks from drifting seconds apart on step 1."""
        if self.pp_group is None or self.world_size < 2:
            return
        dev = self.device if isinstance(self.device, torch.device) else torch.device(self.device)
        ops_, keep = [], []
        for delta in (+1, -1):
            peer_rank = self.rank + delta
            if 0 <= peer_rank < self.world_size:
                peer = dist.get_global_rank(self.pp_group, peer_rank)
                s, r = torch.ones(8, device=dev), torch.empty(8, device=dev)
                keep += [s, r]
                ops_ += [dist.P2POp(dist.isend, s, peer, self.pp_group), dist.P2POp(dist.irecv, r, peer, self.pp_group)]
        for w in dist.batch_isend_irecv(ops_):
            w.wait()
        if dev.type == "cuda":
            torch.cuda.synchronize()

    # ------------------------------------------------------------- loss helpers
    def clm_loss(self, logits: torch.Tensor, labels: torch.Tensor):
        """(sum of token losses, number of supervised tokens) from full logits."""
        V = logits.shape[-1]
        if self.shift_labels:
            logits = logits[:, :-1, :]
            labels = labels[:, 1:]
        return ops.cross_entropy_sum(logits.reshape(-1, V), labels.reshape(-1), -100)

    # --------------------------------------------------------- gradient epilogue
    def sync_auxiliary_grads(self) -> None:
        stage = self.stage
        if hasattr(stage, "sync_sequence_parallel_grads"):
            stage.sync_sequence_parallel_grads()
        if hasattr(stage, "sync_tied_weights_grad"):
            stage.sync_tied_weights_grad()

    def clip_and_step(self) -> None:
        from ...optim.utils import clip_grad_norm_global, global_grad_norm, zero_grads
        opt = self.optimizer
        if self.max_grad_norm > 0 and not getattr(opt, "handles_grad_clipping", False):
            if getattr(opt, "accepts_grad_scale", False):
                # the clip coefficient rides into the fused AdamW kernel as a device scalar: no pass over the gradients
                # to scale them (and no host read of the norm)
                params = [p for p in self.model.parameters()]
                norm = global_grad_norm(params, self.pg_manager)
                coef = None if norm is None else (self.max_grad_norm / (norm + 1e-6)).clamp(max=1.0)
                opt.step(grad_scale=coef)
                zero_grads(self.model, opt)
                return
            clip_grad_norm_global(self.model.parameters(), self.max_grad_norm, self.pg_manager)
        opt.step()
        zero_grads(self.model, opt)

    # ----------------------------------------------------------------- public
    def train_step(self, data_loader, tensor_shapes, device, dtype):
        self.model.train()
        return self.schedule.train_step(data_loader, tensor_shapes, device, dtype)

    def _pad_batch_for_sequence_parallel(self, batch, key):
        """A sequence-parallel stage splits the micro-batch over the TP group: a validation batch (other batch size, ragged
        last batch) that is not divisible by tp is padded with rows whose labels are all ``ignore_index`` — they add
        nothing to the loss sum or the token count."""
        from .schedule import unwrap_stage
        stage = unwrap_stage(self.model)
        if not getattr(stage, "sequence_parallel", False) or self.task_type != "clm":
            return batch
        n = dist.get_world_size(stage.tp_group) if getattr(stage, "tp_group", None) is not None else 1
        b = batch[key].shape[0]
        if n <= 1 or b % n == 0:
            return batch
        pad = n - b % n
        out = dict(batch)
        for k, v in batch.items():
            if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == b:
                fill = -100 if k == "labels" else 0
                out[k] = torch.cat([v, torch.full((pad,) + tuple(v.shape[1:]), fill, dtype=v.dtype, device=v.device)])
        return out

    @torch.no_grad()
    def evaluate(self, val_loader, tensor_shapes, device, dtype):
        """Forward-only pipeline over ``val_loader``.

        Returns ``(avg_loss, metric)`` on the last stage (accuracy % for
        classification, perplexity for CLM) and ``(None, None)`` elsewhere.
        ``tensor_shapes`` describes ONE validation batch.
        """
        self.model.eval()
        if self.schedule_type == "1f1b_interleaved":
            key = "input_ids" if self.task_type == "clm" else "images"
            batches = []
            for batch in val_loader:
                if isinstance(batch, (tuple, list)):
                    batch = {"images": batch[0], "labels": batch[1]}
                if "image" in batch:
                    batch = {"images": batch["image"], "labels": batch["label"]}
                if batch[key].shape[0] == tensor_shapes[0]:          # fixed-shape p2p buffers: skip a ragged tail
                    batches.append(batch)
            return self.schedule.evaluate(batches, tensor_shapes, device, dtype)
        loss_acc = torch.zeros((), device=device)
        count = torch.zeros((), device=device)
        correct = torch.zeros((), device=device)
        nb = 0
        for batch in val_loader:
            if isinstance(batch, (tuple, list)):
                batch = {"images": batch[0], "labels": batch[1]}
            if "image" in batch:
                batch = {"images": batch["image"], "labels": batch["label"]}
            key = "input_ids" if self.task_type == "clm" else "images"
            batch = self._pad_batch_for_sequence_parallel(batch, key)
            bsz = batch[key].shape[0]
            shapes = (bsz,) + tuple(tensor_shapes[1:])
            if hasattr(self.model, "get_tensor_shapes"):
                try:
                    shapes = tuple(self.model.get_tensor_shapes(bsz))
                except (RuntimeError, AttributeError):      # stage without static shapes: keep the caller's
                    pass
            x = pipeline_communicate("recv_forward", self.pp_group, self.rank, device, dtype,
                                     self.is_first_stage, self.is_last_stage, shapes=shapes)
            kwargs = {}
            labels = None
            if self.is_last_stage:
                labels = batch["labels"].to(device)
                if self.fused_loss:
                    kwargs["labels"] = labels
            if self.is_first_stage:
                inp = batch[key].to(device)
                if self.task_type != "clm" and self.compute_dtype is not None:
                    inp = inp.to(self.compute_dtype)
                out = self.model.forward(inp, **kwargs)
            else:
                out = self.model.forward(x, **kwargs)
            if not self.is_last_stage:
                pipeline_communicate("send_forward", self.pp_group, self.rank, device, dtype,
                                     self.is_first_stage, self.is_last_stage, tensor=out, shapes=shapes)
                continue
            nb += 1
            if self.task_type == "clm":
                ls, n = out if self.fused_loss else self.clm_loss(out, labels)
                loss_acc += ls.float()
                count += n.float()
            else:
                loss_acc += self.criterion(out.float(), labels).float()
                correct += (out.argmax(dim=1) == labels).sum().float()
                count += float(labels.shape[0])
        drain_pipeline_sends()
        if not self.is_last_stage:
            return None, None
        l, c, n = torch.stack([loss_acc, correct, count]).tolist()
        if self.task_type == "clm":
            avg = l / n if n > 0 else 0.0
            return avg, (math.exp(avg) if avg < 20 else float("inf"))
        return l / max(nb, 1), (100.0 * c / n if n > 0 else 0.0)
