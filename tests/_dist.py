"""Spawn helper: run ``fn(rank, world_size, *args)`` in ``world_size`` processes
with a gloo (CPU) or nccl (GPU) process group on 127.0.0.1."""
import os
import socket
import sys
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, backend, fn, args, errq):
    try:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                          WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
        torch.manual_seed(1234)
        torch.set_num_threads(1)
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(rank)
            kw["device_id"] = torch.device("cuda", rank)
        dist.init_process_group(backend, rank=rank, world_size=world_size, **kw)
        fn(rank, world_size, *args)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put((rank, traceback.format_exc()))
        raise


def run_distributed(fn, world_size: int, *args, backend: str = "gloo", timeout: int = 240):
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world_size, port, backend, fn, args, errq))
             for r in range(world_size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    hung = [p for p in procs if p.is_alive()]
    for p in hung:
        p.kill()
    if errs:
        raise AssertionError("distributed worker failed:\n" + "\n".join(f"[rank {r}]\n{tb}" for r, tb in errs))
    if hung:
        raise AssertionError(f"{len(hung)} distributed workers hung")
    bad = [p.exitcode for p in procs if p.exitcode != 0]
    assert not bad, f"worker exit codes {bad}"
