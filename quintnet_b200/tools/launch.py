"""Local launcher: run an example under ``torchrun`` on this node, stream its output and parse the epoch
metrics it prints — the job the reference's Modal cloud launchers do (``train_modal_run.py:101-148`` regex-parse
``"Train Loss: x | Train Acc: y%"``), without the cloud dependency.

    python -m quintnet_b200.tools.launch --nproc 8 -m examples.full_3d -- --config examples/config.yaml --synthetic
"""
from __future__ import annotations

import argparse
import json
import re
import subprocess
import sys
import time

_PATTERNS = {
    "train": re.compile(r"Train Loss:\s*([\d.]+)\s*\|\s*Train (Acc|PPL):\s*([\d.]+)"),
    "val": re.compile(r"Val Loss:\s*([\d.]+)\s*\|\s*Val\s+(Acc|PPL):\s*([\d.]+)"),
    "time": re.compile(r"Total training time:\s*([\d.]+)"),
}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--nproc", type=int, default=1)
    ap.add_argument("--port", type=int, default=29571)
    ap.add_argument("-m", dest="module", required=True)
    ap.add_argument("--timeout", type=int, default=0)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    extra = [x for x in a.rest if x != "--"]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(a.port), "-m", a.module, *extra]
    t0 = time.time()
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    epochs, cur = [], {}
    try:
        for line in proc.stdout:
            sys.stdout.write(line)
            for key, pat in _PATTERNS.items():
                m = pat.search(line)
                if not m:
                    continue
                if key == "time":
                    cur["total_s"] = float(m.group(1))
                else:
                    cur[f"{key}_loss"] = float(m.group(1))
                    cur[f"{key}_{m.group(2).lower()}"] = float(m.group(3))
                    if key == "val":
                        epochs.append(dict(cur))
                        cur = {}
            if a.timeout and time.time() - t0 > a.timeout:
                proc.kill()
                break
    finally:
        rc = proc.wait()
    print(json.dumps({"exit_code": rc, "wall_s": time.time() - t0, "epochs": epochs, **cur}))
    return rc


if __name__ == "__main__":
    sys.exit(main())
