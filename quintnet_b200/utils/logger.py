"""Per-rank log tee (reference ``utils/logger.py:5-41``) + structured JSONL step log."""
from __future__ import annotations

import json
import os
import sys
import time
from typing import Optional


class Logger:
    """Tee a stream into a file."""

    def __init__(self, stream, path: str):
        self.stream = stream
        self.file = open(path, "a", buffering=1)

    def write(self, data):
        self.stream.write(data)
        self.file.write(data)

    def flush(self):
        self.stream.flush()
        self.file.flush()


def setup_rank_logging(log_dir: str = "logs") -> str:
    rank = int(os.environ.get("RANK", 0))
    os.makedirs(log_dir, exist_ok=True)
    path = os.path.join(log_dir, f"rank_{rank}.log")
    sys.stdout = Logger(sys.stdout, path)
    sys.stderr = Logger(sys.stderr, path)
    return path


class StepLogger:
    """One JSON object per line (loss, tokens/s, per-phase ms …); rank 0 only."""

    def __init__(self, path: Optional[str], rank: int = 0):
        self.f = None
        if path and rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            self.f = open(path, "a", buffering=1)

    def log(self, **kw):
        if self.f is not None:
            kw.setdefault("time", time.time())
            self.f.write(json.dumps(kw) + "\n")


def setup_logger(name: str = "quintnet", level: str = "INFO"):
    import logging
    lg = logging.getLogger(name)
    if not lg.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter("%(asctime)s [%(levelname)s] %(message)s"))
        lg.addHandler(h)
    lg.setLevel(getattr(logging, level.upper(), logging.INFO))
    return lg


def log_rank_0(msg: str, logger=None) -> None:
    if int(os.environ.get("RANK", 0)) == 0:
        (logger.info if logger is not None else print)(msg)
