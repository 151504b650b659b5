"""Pipeline schedules: all-forward-all-backward (AFAB) and 1F1B.

Same micro-batch choreography as the reference
(``parallelism/pipeline_parallel/schedule.py:74-516``: AFAB = M forwards then M
FIFO backwards; 1F1B = ``min(pp-1-rank, M)`` warm-up forwards, steady
one-forward-one-backward with batched bidirectional p2p, cool-down backwards;
clip + ``optimizer.step()`` + ``zero_grad()`` inside the schedule; returns
``(loss, metric)`` on the last stage and ``(None, None)`` elsewhere).

Re-designed for a GPU that is ~50x faster than the host round-trips the
reference pays per micro-batch:

* loss and metrics accumulate in device tensors; ONE host read per optimizer step;
* no device-wide synchronisation anywhere; sends stay in flight;
* the loss is scaled by 1/M (``loss_scale='mean'``; ``'legacy_sum'`` reproduces
  the reference's summed-gradient behaviour);
* CLM labels are shifted (next-token prediction) unless ``shift_labels=False``;
* data-parallel gradient reduction is triggered only by the LAST micro-batch's
  backward (``require_backward_grad_sync``);
* tied wte/lm_head gradients and sequence-parallel replicated-parameter
  gradients are reduced BEFORE the optimizer step;
* gradient clipping uses the global norm over (tp, pp).
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod
from typing import List, Optional

import torch

from ...core.communication import (bidirectional_pipeline_communicate, drain_pipeline_sends,
                                   pipeline_communicate, post_recv, post_send)


def unwrap_stage(model):
    """DP -> PP wrapper -> stage module (attribute sniffing as in the reference trainers)."""
    m = model
    if hasattr(m, "bucket_manager") and hasattr(m, "model"):
        m = m.model
    if hasattr(m, "local_module"):
        m = m.local_module
    return m


def _set_grad_sync(model, flag: bool) -> None:
    if hasattr(model, "require_backward_grad_sync"):
        model.require_backward_grad_sync = flag


class PipelineSchedule(ABC):
    def __init__(self, trainer):
        self.trainer = trainer
        self._loss_acc = None
        self._count_acc = None
        self._correct_acc = None
        self._nmb = 0

    @abstractmethod
    def train_step(self, data_loader, tensor_shapes, device, dtype): ...

    # ------------------------------------------------------------ p2p sugar
    def _comm(self, op, device, dtype, shapes, tensor=None):
        t = self.trainer
        return pipeline_communicate(operation=op, pp_group=t.pp_group, pp_rank=t.rank, device=device,
                                    dtype=dtype, is_first_stage=t.is_first_stage,
                                    is_last_stage=t.is_last_stage, tensor=tensor, shapes=shapes)

    def _defer_forward_recv(self) -> bool:
        """True when this stage runs sequence-parallel tensor parallelism (the fused TP kernels): see the 1F1B steady
        state.  ``QN_PP_PREPOST_FWD=1`` / ``0`` forces pre-posting on / off (A/B timing)."""
        import os
        v = os.environ.get("QN_PP_PREPOST_FWD")
        if v in ("0", "1"):
            return v == "0"
        stage = unwrap_stage(self.trainer.model)
        return bool(getattr(stage, "sequence_parallel", False))      # the stage sets it only when its tp group is > 1

    def _bidir(self, op, send, device, dtype, shapes):
        t = self.trainer
        return bidirectional_pipeline_communicate(operation=op, pp_group=t.pp_group, pp_rank=t.rank,
                                                  send_tensor=send, recv_shapes=shapes, device=device,
                                                  dtype=dtype, is_first_stage=t.is_first_stage,
                                                  is_last_stage=t.is_last_stage)

    # ------------------------------------------------------------- one step
    def _begin(self, device):
        self._loss_acc = torch.zeros((), device=device, dtype=torch.float32)
        self._count_acc = torch.zeros((), device=device, dtype=torch.float32)
        self._correct_acc = torch.zeros((), device=device, dtype=torch.float32)
        self._nmb = 0
        self._bwd_done = 0

    def _forward_step(self, input_tensor, data_loader, device, num_microbatches: int):
        t = self.trainer
        batch = None
        if t.is_first_stage or t.is_last_stage:
            batch = next(data_loader)
        kwargs = {}
        labels = None
        if t.is_last_stage:
            labels = batch["labels"].to(device, non_blocking=True)
            if t.task_type == "clm" and t.fused_loss:
                kwargs["labels"] = labels
        if t.is_first_stage:
            key = "input_ids" if t.task_type == "clm" else "images"
            x = batch[key].to(device, non_blocking=True)
            if t.task_type != "clm" and x.is_floating_point() and t.compute_dtype is not None:
                x = x.to(t.compute_dtype)
            out = t.model.forward(x, **kwargs)
        else:
            out = t.model.forward(input_tensor, **kwargs)
        if not t.is_last_stage:
            return out

        scale = 1.0 if t.loss_scale == "legacy_sum" else 1.0 / num_microbatches
        if t.task_type == "clm":
            if t.fused_loss:
                loss_sum, n_valid = out
            else:
                loss_sum, n_valid = t.clm_loss(out, labels)
            n_valid = n_valid.to(torch.float32)
            loss = loss_sum / n_valid.clamp(min=1.0)
            self._loss_acc += loss_sum.detach().float()
            self._count_acc += n_valid
        else:
            loss = t.criterion(out.float(), labels)
            self._loss_acc += loss.detach().float()
            self._correct_acc += (out.argmax(dim=1) == labels).sum().float()
            self._count_acc += float(labels.shape[0])
        self._nmb += 1
        return loss * scale

    def _backward_step(self, input_tensor, output_tensor, output_grad, num_microbatches: int):
        t = self.trainer
        self._bwd_done += 1
        _set_grad_sync(t.model, self._bwd_done == num_microbatches)
        return t.model.backward(input_tensor, output_tensor, output_grad)

    def _finish_step(self):
        t = self.trainer
        drain_pipeline_sends()
        _set_grad_sync(t.model, True)
        if hasattr(t.model, "finish_gradient_sync"):
            t.model.finish_gradient_sync()
        t.sync_auxiliary_grads()
        if t.optimizer is not None:
            t.clip_and_step()
        if not t.is_last_stage:
            return None, None
        # single host read per optimizer step
        if t.task_type == "clm":
            loss_sum, count = torch.stack([self._loss_acc, self._count_acc]).tolist()
            avg = loss_sum / count if count > 0 else 0.0
            return avg, (math.exp(avg) if avg < 20 else float("inf"))
        loss_sum, correct, count = torch.stack(
            [self._loss_acc, self._correct_acc, self._count_acc]).tolist()
        return loss_sum / max(self._nmb, 1), (100.0 * correct / count if count > 0 else 0.0)


class AllFwdAllBwdSchedule(PipelineSchedule):
    def train_step(self, data_loader, tensor_shapes, device, dtype):
        M = data_loader.grad_acc_steps
        self._begin(device)
        inputs: List[Optional[torch.Tensor]] = []
        outputs: List[torch.Tensor] = []
        for _ in range(M):
            x = self._comm("recv_forward", device, dtype, tensor_shapes)
            y = self._forward_step(x, data_loader, device, M)
            self._comm("send_forward", device, dtype, tensor_shapes, tensor=y)
            inputs.append(x)
            outputs.append(y)
        for _ in range(M):
            g = self._comm("recv_backward", device, dtype, tensor_shapes)
            x, y = inputs.pop(0), outputs.pop(0)
            gx = self._backward_step(x, y, g, M)
            self._comm("send_backward", device, dtype, tensor_shapes, tensor=gx)
        return self._finish_step()


class OneFOneBSchedule(PipelineSchedule):
    def train_step(self, data_loader, tensor_shapes, device, dtype):
        t = self.trainer
        M = data_loader.grad_acc_steps
        self._begin(device)
        n_warm = min(t.world_size - t.rank - 1, M)
        n_steady = M - n_warm
        inputs: List[Optional[torch.Tensor]] = []
        outputs: List[torch.Tensor] = []

        for _ in range(n_warm):
            x = self._comm("recv_forward", device, dtype, tensor_shapes)
            y = self._forward_step(x, data_loader, device, M)
            self._comm("send_forward", device, dtype, tensor_shapes, tensor=y)
            inputs.append(x)
            outputs.append(y)

        # steady state with PRE-POSTED receives: the gradient of the oldest in-flight micro-batch and the activations of
        # the NEXT micro-batch are both posted before the backward runs, so they arrive on the communication streams
        # (one communicator per direction, see core/communication.py) while this stage computes; sends stay in flight.
        kw = dict(pp_group=t.pp_group, pp_rank=t.rank, is_first_stage=t.is_first_stage, is_last_stage=t.is_last_stage)
        # A posted receive is a RESIDENT NCCL kernel until its data arrives.  The fused TP kernels (AG->GEMM / GEMM->RS)
        # are persistent grids that need every one of their CTAs to run before the tp peer can finish, so an NCCL CTA
        # parked on an SM for the length of a backward pass can stall them until the neighbouring stage delivers.  With
        # sequence-parallel TP the next activation is therefore posted AFTER the backward (its 16 MB transfer is ~50 us
        # of a ~10 ms backward: nothing to hide); without TP it stays pre-posted.
        defer_fwd_recv = self._defer_forward_recv()
        x = self._comm("recv_forward", device, dtype, tensor_shapes) if n_steady > 0 else None
        for i in range(n_steady):
            last = i == n_steady - 1
            y = self._forward_step(x, data_loader, device, M)
            if not t.is_last_stage:
                post_send("forward", tensor=y, **kw)
            g_pending = post_recv("backward", device=device, dtype=dtype, shapes=tensor_shapes, **kw)
            x_pending = None
            if not last and not defer_fwd_recv:
                x_pending = post_recv("forward", device=device, dtype=dtype, shapes=tensor_shapes, **kw)
            inputs.append(x)
            outputs.append(y)
            bx, by = inputs.pop(0), outputs.pop(0)
            gx = self._backward_step(bx, by, g_pending.wait(), M)
            if not t.is_first_stage:
                post_send("backward", tensor=gx, **kw)
            if not last and defer_fwd_recv:
                x_pending = post_recv("forward", device=device, dtype=dtype, shapes=tensor_shapes, **kw)
            x = x_pending.wait() if x_pending is not None else None

        for _ in range(n_warm):
            bx, by = inputs.pop(0), outputs.pop(0)
            g = self._comm("recv_backward", device, dtype, tensor_shapes)
            gx = self._backward_step(bx, by, g, M)
            self._comm("send_backward", device, dtype, tensor_shapes, tensor=gx)
        return self._finish_step()
