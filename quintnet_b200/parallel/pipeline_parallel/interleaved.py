"""Interleaved 1F1B ("virtual pipeline stages").

Not in the reference (its schedules are AFAB and plain 1F1B, ``parallelism/pipeline_parallel/schedule.py:82-516``);
added because at the benchmark's shape (pp = 2, 8 micro-batches) the 1F1B bubble is 1/9 of the step: with V model
chunks per rank the pipeline has S = pp·V stages (stage s lives on rank s mod pp) and the bubble shrinks by V.

Design: the schedule is a TABLE, computed once per (pp, V, M) by every rank identically.

1. Each rank's unit order is the Megatron-LM interleaved order (warm-up forwards, then one-forward-one-backward,
   then cool-down backwards; micro-batches advance in groups of pp per chunk).
2. A tick simulation assigns every unit a global tick: a unit runs at the first tick at which the unit it depends
   on (previous stage's forward / next stage's backward of the same micro-batch) finished at an EARLIER tick.
3. Whatever a rank produces in tick t is sent in its tick-t communication batch, and the consumer posts the
   matching receive in ITS tick-t batch.  Every rank therefore issues one ``batch_isend_irecv`` per tick, in tick
   order, and a send and its receive always sit in the same tick on both sides — the NCCL group semantics make
   that deadlock-free by construction (no hand-paired send/recv special cases), and gloo behaves identically.

Receives are waited for only when the unit that consumes them runs, so transfers overlap with compute.
"""
from __future__ import annotations

import functools
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from .schedule import PipelineSchedule, _set_grad_sync

Unit = Tuple[str, int, int]          # ('F' | 'B', chunk v, micro-batch m)


def unit_order(pp: int, V: int, M: int, rank: int, forward_only: bool = False) -> List[Unit]:
    """Megatron-LM interleaved 1F1B order of one rank (requires M % pp == 0)."""
    total = M * V

    def f_unit(k):
        return (k // pp) % V, (k // (pp * V)) * pp + (k % pp)

    def b_unit(k):
        return V - 1 - ((k // pp) % V), (k // (pp * V)) * pp + (k % pp)

    if forward_only:
        return [("F",) + f_unit(k) for k in range(total)]
    warm = min((pp - rank - 1) * 2 + (V - 1) * pp, total)
    order: List[Unit] = [("F",) + f_unit(k) for k in range(warm)]
    fk, bk = warm, 0
    while fk < total:
        order.append(("F",) + f_unit(fk)); fk += 1
        order.append(("B",) + b_unit(bk)); bk += 1
    while bk < total:
        order.append(("B",) + b_unit(bk)); bk += 1
    return order


class Tick:
    __slots__ = ("unit", "sends", "recvs")

    def __init__(self):
        self.unit: Optional[Unit] = None
        self.sends: List[Tuple[int, Unit]] = []      # (peer pp-rank, unit that CONSUMES the tensor)
        self.recvs: List[Tuple[int, Unit]] = []      # (peer pp-rank, local unit that consumes it)


@functools.lru_cache(maxsize=32)
def build_table(pp: int, V: int, M: int, forward_only: bool = False) -> Tuple[Tuple[Tick, ...], ...]:
    """Per-rank tuple of ticks (all ranks have the same number of ticks)."""
    if pp < 2 or V < 1:
        raise ValueError("interleaved schedule needs pp >= 2")
    if M % pp:
        raise ValueError(f"interleaved 1F1B needs the number of micro-batches ({M}) to be a multiple of pp ({pp})")
    S = pp * V
    orders = [unit_order(pp, V, M, r, forward_only) for r in range(pp)]
    idx = [0] * pp
    done: Dict[Tuple[str, int, int], int] = {}            # (kind, stage, mb) -> tick it ran in
    table: List[List[Tick]] = [[] for _ in range(pp)]
    t = 0
    while any(idx[r] < len(orders[r]) for r in range(pp)):
        ran = []
        for r in range(pp):
            tick = Tick()
            table[r].append(tick)
            if idx[r] >= len(orders[r]):
                continue
            kind, v, m = orders[r][idx[r]]
            s = v * pp + r
            if kind == "F":
                dep = ("F", s - 1, m) if s > 0 else None
            else:
                dep = ("B", s + 1, m) if s < S - 1 else ("F", s, m)
            if dep is not None and not (dep in done and done[dep] < t):
                continue
            tick.unit = (kind, v, m)
            ran.append((r, kind, s, m))
            idx[r] += 1
        if not ran:
            raise RuntimeError(f"interleaved schedule deadlocked at tick {t} (pp={pp}, V={V}, M={M})")
        for r, kind, s, m in ran:
            done[(kind, s, m)] = t
            if kind == "F" and s < S - 1:
                dst, cons = (s + 1) % pp, ("F", (s + 1) // pp, m)
            elif kind == "B" and s > 0:
                dst, cons = (s - 1) % pp, ("B", (s - 1) // pp, m)
            else:
                continue
            table[r][t].sends.append((dst, cons))
            table[dst][t].recvs.append((r, cons))
        t += 1
    return tuple(tuple(ticks) for ticks in table)


def bubble_fraction(pp: int, V: int, M: int, tf: float = 1.0, tb: float = 2.0) -> float:
    """Idle fraction of the critical rank under the table, with forward / backward unit costs tf / tb (a unit of a
    V-chunk model costs 1/V of the un-chunked stage)."""
    table = build_table(pp, V, M)
    ready = {}
    clock = [0.0] * pp
    S = pp * V
    for t in range(len(table[0])):
        for r in range(pp):
            u = table[r][t].unit
            if u is None:
                continue
            kind, v, m = u
            s = v * pp + r
            dep = (("F", s - 1, m) if s > 0 else None) if kind == "F" else (("B", s + 1, m) if s < S - 1 else None)
            start = max(clock[r], ready.get(dep, 0.0)) if dep else clock[r]
            clock[r] = start + (tf if kind == "F" else tb) / V
            ready[(kind, s, m)] = clock[r]
    busy = M * (tf + tb)
    return 1.0 - busy / max(clock)


# ------------------------------------------------------------------------------------------------ modules
class ChunkedStage(nn.Module):
    """The V model chunks of one rank behind the attribute surface the trainers expect from a stage."""

    def __init__(self, chunks: List[nn.Module]):
        super().__init__()
        self.chunks = nn.ModuleList(chunks)

    # attributes the trainers set / read on a stage
    @property
    def supports_fused_loss(self) -> bool:
        return bool(getattr(self.chunks[-1], "supports_fused_loss", False))

    @property
    def shift_labels(self):
        return getattr(self.chunks[-1], "shift_labels", True)

    @shift_labels.setter
    def shift_labels(self, v):
        for c in self.chunks:
            if hasattr(c, "shift_labels"):
                c.shift_labels = v

    @property
    def seq_len(self):
        return getattr(self.chunks[0], "seq_len", None)

    @seq_len.setter
    def seq_len(self, v):
        for c in self.chunks:
            if hasattr(c, "seq_len"):
                c.seq_len = v

    def get_tensor_shapes(self, batch_size: int):
        return self.chunks[0].get_tensor_shapes(batch_size)

    def sync_sequence_parallel_grads(self):
        for c in self.chunks:
            if hasattr(c, "sync_sequence_parallel_grads"):
                c.sync_sequence_parallel_grads()

    def sync_tied_weights_grad(self):
        for c in self.chunks:
            if hasattr(c, "sync_tied_weights_grad"):
                c.sync_tied_weights_grad()

    def forward(self, x, chunk: int = 0, **kw):
        return self.chunks[chunk](x, **kw) if kw else self.chunks[chunk](x)


class VirtualPipelineWrapper(nn.Module):
    """Pipeline wrapper of a rank that owns V chunks (virtual stages ``v * pp + rank``).  Same attribute contract as
    :class:`PipelineParallelWrapper` (``local_module``, ``rank``, ``world_size``, ``group``, ``is_first_stage``,
    ``is_last_stage``, ``get_tensor_shapes``, ``backward``); ``forward`` takes the chunk index."""

    def __init__(self, chunks: List[nn.Module], device_mesh, rank: int, pp_group, pp_size: int, device):
        super().__init__()
        self.device_mesh = device_mesh
        self.rank, self.world_size, self.group = rank, pp_size, pp_group
        self.is_first_stage = rank == 0                   # owns virtual stage 0 (embedding)
        self.is_last_stage = rank == pp_size - 1          # owns the last virtual stage (lm head / loss)
        self.device = device
        self.num_chunks = len(chunks)
        self.local_module = ChunkedStage([c.to(device) for c in chunks])

    def get_tensor_shapes(self, batch_size: int):
        return tuple(self.local_module.get_tensor_shapes(batch_size))

    def forward(self, x, chunk: int = 0, **kwargs):
        return self.local_module(x, chunk=chunk, **kwargs)

    @staticmethod
    def backward(input_tensor, output_tensor, output_tensor_grad):
        if input_tensor is not None and input_tensor.requires_grad:
            input_tensor.retain_grad()
        if output_tensor_grad is None and output_tensor.numel() == 1:
            torch.autograd.backward(output_tensor)
        else:
            if output_tensor_grad is None:
                output_tensor_grad = torch.ones_like(output_tensor)
            torch.autograd.backward(output_tensor, grad_tensors=output_tensor_grad)
        grad = None
        if input_tensor is not None:
            grad = input_tensor.grad
            input_tensor.grad = None
        return grad


# ------------------------------------------------------------------------------------------------ schedule
class InterleavedOneFOneBSchedule(PipelineSchedule):
    """Table-driven interleaved 1F1B (see the module docstring)."""

    def _pp_wrapper(self):
        m = self.trainer.model
        if hasattr(m, "bucket_manager") and hasattr(m, "model"):
            m = m.model
        return m

    def _run(self, data_loader, tensor_shapes, device, dtype, M: int, forward_only: bool):
        t = self.trainer
        wrap = self._pp_wrapper()
        pp, V, rank = t.world_size, wrap.num_chunks, t.rank
        table = build_table(pp, V, M, forward_only)[rank]
        S = pp * V
        inbox: Dict[Unit, Tuple[torch.Tensor, object]] = {}
        saved_in: Dict[Tuple[int, int], Optional[torch.Tensor]] = {}
        saved_out: Dict[Tuple[int, int], torch.Tensor] = {}
        pending_sends = []
        bwd_total = M * V

        def take(u: Unit):
            buf, work = inbox.pop(u)
            work.wait()
            return buf

        for tick in table:
            produced = None
            if tick.unit is not None:
                kind, v, m = tick.unit
                s = v * pp + rank
                if kind == "F":
                    x = take(tick.unit) if s > 0 else None
                    y = self._forward_unit(v, s, S, x, data_loader, device, M)
                    if not forward_only:
                        saved_in[(v, m)], saved_out[(v, m)] = x, y
                    produced = y if s < S - 1 else None
                else:
                    g = take(tick.unit) if s < S - 1 else None
                    x, y = saved_in.pop((v, m)), saved_out.pop((v, m))
                    self._bwd_done += 1
                    # gradient sync only on the globally LAST backward unit: DataParallel's end-of-backward callback
                    # reduces every bucket that has not been launched — also those of chunks whose gradients are
                    # still accumulating — so enabling it per chunk would all-reduce the whole model V times (and
                    # scale the gradients by dp on every extra pass under ReductionStrategy.SUM)
                    _set_grad_sync(t.model, self._bwd_done == bwd_total)
                    produced = wrap.backward(x, y, g)
                    if s == 0:
                        produced = None
            ops = []
            for peer, _cons in tick.sends:
                buf = produced.detach().contiguous()
                ops.append(dist.P2POp(dist.isend, buf, dist.get_global_rank(t.pp_group, peer), t.pp_group))
                pending_sends.append(buf)
            recv_bufs = []
            for peer, cons in tick.recvs:
                buf = torch.empty(tensor_shapes, device=device, dtype=dtype,
                                  requires_grad=(cons[0] == "F" and not forward_only))
                ops.append(dist.P2POp(dist.irecv, buf, dist.get_global_rank(t.pp_group, peer), t.pp_group))
                recv_bufs.append((cons, buf))
            if ops:
                works = dist.batch_isend_irecv(ops)
                n_send = len(tick.sends)
                for w in works[:n_send]:
                    pending_sends.append(w)
                for i, (cons, buf) in enumerate(recv_bufs):
                    # some backends return ONE work for the whole batch
                    inbox[cons] = (buf, works[min(n_send + i, len(works) - 1)])
        for w in pending_sends:
            if hasattr(w, "wait"):
                w.wait()
        assert not inbox and not saved_in, "interleaved schedule left unconsumed tensors"
        assert forward_only or self._bwd_done == bwd_total

    def _forward_unit(self, v: int, s: int, S: int, x, data_loader, device, M: int):
        """Forward of chunk v (global stage s).  Stage 0 reads the inputs, stage S-1 the labels and returns the
        scaled loss (accumulating the metrics on the device)."""
        t = self.trainer
        wrap = self._pp_wrapper()
        first, last = s == 0, s == S - 1
        kwargs = {}
        labels = None
        batch = next(data_loader) if (first or last) else None
        if last:
            labels = batch["labels"].to(device, non_blocking=True)
            if t.task_type == "clm" and t.fused_loss:
                kwargs["labels"] = labels
        if first:
            key = "input_ids" if t.task_type == "clm" else "images"
            x = batch[key].to(device, non_blocking=True)
        out = wrap.forward(x, chunk=v, **kwargs)
        if not last:
            return out
        scale = 1.0 if t.loss_scale == "legacy_sum" else 1.0 / M
        if t.task_type == "clm":
            loss_sum, n_valid = out if t.fused_loss else t.clm_loss(out, labels)
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

    def train_step(self, data_loader, tensor_shapes, device, dtype):
        M = data_loader.grad_acc_steps
        self._begin(device)
        self._run(data_loader, tuple(tensor_shapes), device, dtype, M, forward_only=False)
        return self._finish_step()

    @torch.no_grad()
    def evaluate(self, batches, tensor_shapes, device, dtype):
        """Forward-only pass over ``batches`` (an iterator of dict batches; its length must be a multiple of pp —
        the tail that does not fill a group is dropped).  Returns ``(avg_loss, metric)`` on the last rank."""
        t = self.trainer
        batches = list(batches)
        n = (len(batches) // t.world_size) * t.world_size
        self._begin(device)
        if n:
            it = iter(batches[:n])
            self._run(it, tuple(tensor_shapes), device, dtype, n, forward_only=True)
        if not t.is_last_stage:
            return None, None
        if t.task_type == "clm":
            loss_sum, count = torch.stack([self._loss_acc, self._count_acc]).tolist()
            avg = loss_sum / count if count > 0 else 0.0
            return avg, (math.exp(avg) if avg < 20 else float("inf"))
        loss_sum, correct, count = torch.stack([self._loss_acc, self._correct_acc, self._count_acc]).tolist()
        return loss_sum / max(self._nmb, 1), (100.0 * correct / count if count > 0 else 0.0)
