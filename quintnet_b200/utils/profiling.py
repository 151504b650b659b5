"""Timing / memory profiling (reference ``utils/profiling.py:11-27`` are broken
stubs).  Device timing uses CUDA events on the current stream; multi-rank
numbers are reduced with MAX (the slowest rank defines the step)."""
from __future__ import annotations

import contextlib
import os
import time
from collections import defaultdict
from typing import Dict, Optional

import torch
import torch.distributed as dist


class DeviceTimer:
    """Accumulates named phase durations.  CUDA: event pairs resolved lazily at
    ``summary()`` (no sync in the hot loop); CPU: ``perf_counter``."""

    def __init__(self, enabled: bool = True):
        self.enabled = enabled
        self._events = defaultdict(list)
        self._cpu = defaultdict(float)
        self._counts = defaultdict(int)

    @contextlib.contextmanager
    def phase(self, name: str):
        if not self.enabled:
            yield
            return
        if torch.cuda.is_available():
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            try:
                yield
            finally:
                e.record()
                self._events[name].append((s, e))
                self._counts[name] += 1
        else:
            t0 = time.perf_counter()
            try:
                yield
            finally:
                self._cpu[name] += (time.perf_counter() - t0) * 1e3
                self._counts[name] += 1

    def summary(self, reduce_max: bool = False) -> Dict[str, float]:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        out = dict(self._cpu)
        for name, pairs in self._events.items():
            out[name] = out.get(name, 0.0) + sum(s.elapsed_time(e) for s, e in pairs)
        if reduce_max and dist.is_initialized() and out:
            keys = sorted(out)
            dev = "cuda" if (torch.cuda.is_available() and dist.get_backend() == "nccl") else "cpu"
            t = torch.tensor([out[k] for k in keys], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out = dict(zip(keys, t.tolist()))
        return out

    def reset(self):
        self._events.clear()
        self._cpu.clear()
        self._counts.clear()


@contextlib.contextmanager
def profile_time(name: str = "Operation", rank: int = 0, sink: Optional[dict] = None, verbose: bool = False):
    """Wall-time (device-synchronised) of a block in milliseconds; printed (``verbose``) only on ``rank``
    (reference signature ``(name, rank)``, ``utils/profiling.py:12`` — a stub there)."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        yield
    finally:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        if sink is not None:
            sink[name] = ms
        if verbose and int(os.environ.get("RANK", 0)) == int(rank):
            print(f"[profile_time] {name}: {ms:.3f} ms")


@contextlib.contextmanager
def profile_memory(device: Optional[torch.device] = None, name: str = "block", sink: Optional[dict] = None,
                   verbose: bool = False):
    """Peak CUDA memory delta of a block in MiB on ``device`` (current device by default; 0 on CPU).  Reference
    signature ``(device)``, ``utils/profiling.py:21`` — a stub there."""
    cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == "cuda")
    if cuda:
        torch.cuda.synchronize(device)
        torch.cuda.reset_peak_memory_stats(device)
        base = torch.cuda.memory_allocated(device)
    try:
        yield
    finally:
        peak = 0.0
        if cuda:
            torch.cuda.synchronize(device)
            peak = (torch.cuda.max_memory_allocated(device) - base) / 2 ** 20
        if sink is not None:
            sink[name] = peak
        if verbose:
            print(f"[profile_memory] {name}: peak +{peak:.1f} MiB")


@contextlib.contextmanager
def nvtx_range(name: str):
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield
