"""Per-rank log tee (reference ``utils/logger.py:5-41``) + structured JSONL step log."""
from __future__ import annotations

import json
import os
import sys
import time
from typing import Optional


class Logger:
    """Tee a stream into a file (reference signature ``Logger(filename, stream)``, ``utils/logger.py:5-21``)."""

    def __init__(self, filename: str, stream=None):
        self.stream = self.terminal = stream if stream is not None else sys.stdout
        self.file = self.log = open(filename, "a", buffering=1)

    def write(self, message):
        self.stream.write(message)
        self.file.write(message)

    def flush(self):
        self.stream.flush()
        self.file.flush()


def setup_rank_logging(log_dir: str = "logs") -> str:
    rank = int(os.environ.get("RANK", 0))
    os.makedirs(log_dir, exist_ok=True)
    path = os.path.join(log_dir, f"rank_{rank}.log")
    sys.stdout = Logger(path, sys.stdout)
    sys.stderr = Logger(path, sys.stderr)
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


def setup_logger(name: str = "quintnet", level="INFO", log_file: Optional[str] = None):
    """Rank-tagged ``logging.Logger`` (reference signature ``(name, level, log_file)``, ``utils/logging.py:9`` — a stub
    there).  ``level`` is a ``logging`` int or its name; ``log_file`` adds a file handler."""
    import logging
    lg = logging.getLogger(name)
    fmt = logging.Formatter(f"%(asctime)s [rank {os.environ.get('RANK', 0)}] [%(levelname)s] %(message)s")
    if not lg.handlers:
        h = logging.StreamHandler()
        h.setFormatter(fmt)
        lg.addHandler(h)
    if log_file and not any(getattr(h, "baseFilename", None) == os.path.abspath(log_file) for h in lg.handlers):
        os.makedirs(os.path.dirname(os.path.abspath(log_file)), exist_ok=True)
        fh = logging.FileHandler(log_file)
        fh.setFormatter(fmt)
        lg.addHandler(fh)
    lg.setLevel(level if isinstance(level, int) else getattr(logging, str(level).upper(), logging.INFO))
    return lg


def log_rank_0(message: str, level=None, logger=None) -> None:
    """Emit ``message`` on global rank 0 only (reference signature ``(message, level)``, ``utils/logging.py:22``)."""
    import logging
    if int(os.environ.get("RANK", 0)) != 0:
        return
    if isinstance(level, logging.Logger):          # older call style: log_rank_0(msg, logger)
        logger, level = level, None
    if logger is None:
        print(message)
    else:
        logger.log(level if isinstance(level, int) else logging.INFO, message)
