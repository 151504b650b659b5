"""Job workflow CLI: stage model / dataset → train → list / download checkpoints → merge → test.

The reference drives this workflow through three Modal cloud launchers — ``train_modal_run.py:38-148`` (ViT 3-D run on
8 GPUs, epoch metrics regex-parsed from the log), ``gpt2_train_modal_run.py:73-1097`` (``upload_model``,
``upload_dataset``, ``prepare_dataset``, ``run_gpt2_training``, ``list_checkpoints``, ``download_checkpoints``,
``test_weight_shapes``, ``merge_and_test_model``) and ``test_modal_run.py:38-76`` (run a test file on the box).  Their
Modal *volume* is the only state they share.  Here the same steps work against a **volume directory** (a path on a
shared file system, or ``user@host:/path`` which is synchronised with rsync over ssh), so the workflow runs on a
bare-metal B200 node, a Slurm allocation or a laptop alike::

    python -m quintnet_b200.tools.cluster --volume /mnt/qn upload-model  --src gpt2.safetensors
    python -m quintnet_b200.tools.cluster --volume /mnt/qn upload-dataset --src cnn_dailymail/
    python -m quintnet_b200.tools.cluster --volume /mnt/qn prepare-dataset --name cnn_dailymail --max-samples 10000
    python -m quintnet_b200.tools.cluster --volume /mnt/qn train --workload gpt2 --nproc 8 --run exp1 --config examples/gpt2_config.yaml
    python -m quintnet_b200.tools.cluster --volume /mnt/qn list-checkpoints
    python -m quintnet_b200.tools.cluster --volume /mnt/qn merge-and-test --run exp1 --prefix final_model
    python -m quintnet_b200.tools.cluster --volume /mnt/qn download-checkpoints --run exp1 --dst ./ckpt
    python -m quintnet_b200.tools.cluster --volume /mnt/qn test -- tests/test_mesh.py -q
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import time
from typing import List, Optional

from . import launch as _launch

SUBDIRS = ("models", "datasets", "checkpoints", "logs")


# ----------------------------------------------------------------------------------------------- volume
class Volume:
    """A directory tree shared by every step (the reference's ``modal.Volume``).  Remote specs ``user@host:/path`` are
    mirrored into a local cache with rsync before a step and pushed back after it."""

    def __init__(self, spec: str):
        self.spec = spec
        self.remote = ":" in spec and not os.path.exists(spec.split(":", 1)[0]) and "@" in spec.split(":", 1)[0]
        self.root = os.path.join(os.path.expanduser("~/.cache/quintnet_b200/volumes"),
                                 spec.replace("/", "_").replace(":", "_").replace("@", "_")) if self.remote else spec
        for d in SUBDIRS:
            os.makedirs(os.path.join(self.root, d), exist_ok=True)

    def path(self, *parts: str) -> str:
        return os.path.join(self.root, *parts)

    def pull(self, sub: str = "") -> None:
        if self.remote:
            subprocess.check_call(["rsync", "-a", "--partial", f"{self.spec}/{sub}", self.path(sub)])

    def push(self, sub: str = "") -> None:
        if self.remote:
            subprocess.check_call(["rsync", "-a", "--partial", self.path(sub), f"{self.spec}/{sub}"])


def _copy_in(src: str, dst_dir: str) -> str:
    os.makedirs(dst_dir, exist_ok=True)
    dst = os.path.join(dst_dir, os.path.basename(os.path.normpath(src)))
    if os.path.isdir(src):
        shutil.copytree(src, dst, dirs_exist_ok=True)
    else:
        shutil.copy2(src, dst)
    return dst


def _walk(root: str) -> List[dict]:
    out = []
    for d, _, files in os.walk(root):
        for f in sorted(files):
            p = os.path.join(d, f)
            st = os.stat(p)
            out.append({"path": os.path.relpath(p, root), "bytes": st.st_size,
                        "mtime": time.strftime("%Y-%m-%d %H:%M:%S", time.localtime(st.st_mtime))})
    return sorted(out, key=lambda r: r["path"])


# ------------------------------------------------------------------------------------------------ steps
def upload_model(vol: Volume, src: str) -> str:
    """Stage a HuggingFace-layout GPT-2 checkpoint (``*.safetensors`` file or directory) — ``upload_model``."""
    dst = _copy_in(src, vol.path("models"))
    vol.push("models/")
    return dst


def upload_dataset(vol: Volume, src: str) -> str:
    """Stage a dataset directory (e.g. the CNN/DailyMail ``train.csv`` / ``validation.csv`` pair) — ``upload_dataset``."""
    dst = _copy_in(src, vol.path("datasets"))
    vol.push("datasets/")
    return dst


def prepare_dataset(vol: Volume, name: str, tokenizer: str, max_length: int, max_samples: Optional[int],
                    synthetic: bool, vocab: int = 50257) -> str:
    """Tokenise once and cache the tensors on the volume (``prepare_dataset``): every training rank then memory-maps the
    same file instead of re-tokenising.  ``--synthetic`` writes random tokens of the same shape (no tokenizer files)."""
    import torch
    from ..utils.data import SummarizationCollator, SummarizationDataset, SyntheticTokens
    out = vol.path("datasets", f"{name}_tok{max_length}.pt")
    if synthetic:
        ds = SyntheticTokens(max_samples or 1024, max_length, vocab, seed=0)
        payload = {"train": {"input_ids": ds.input_ids, "labels": ds.labels}}
    else:
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained(tokenizer)
        if tok.pad_token is None:
            tok.pad_token = tok.eos_token
        collate = SummarizationCollator(tok, max_length)
        payload = {}
        for split in ("train", "validation"):
            if not os.path.exists(vol.path("datasets", name, f"{split}.csv")):
                continue
            ds = SummarizationDataset(vol.path("datasets", name), split, max_samples)
            batches = [collate([ds[i] for i in range(j, min(j + 256, len(ds)))]) for j in range(0, len(ds), 256)]
            payload[split] = {k: torch.cat([b[k] for b in batches]) for k in batches[0]}
        if not payload:
            raise FileNotFoundError(f"no train.csv / validation.csv under {vol.path('datasets', name)}")
    torch.save(payload, out)
    vol.push("datasets/")
    return out


def train(vol: Volume, workload: str, nproc: int, run: str, config: Optional[str], extra: List[str], port: int) -> int:
    """``run_3d_parallel_training`` / ``run_gpt2_training``: torchrun the example on this node with its checkpoints
    going to ``<volume>/checkpoints/<run>``; the epoch metrics are parsed from the log like the reference's launchers do."""
    module = {"vit": "examples.full_3d", "gpt2": "examples.gpt2_finetune"}[workload]
    out_dir = vol.path("checkpoints", run)
    os.makedirs(out_dir, exist_ok=True)
    args = ["--nproc", str(nproc), "--port", str(port), "-m", module, "--"]
    if config:
        args += ["--config", config]
    env_before = os.environ.get("QUINTNET_OUTPUT_DIR")
    os.environ["QUINTNET_OUTPUT_DIR"] = out_dir          # the examples write their shards here when set
    try:
        rc = _launch.main(args + list(extra))
    finally:
        if env_before is None:
            os.environ.pop("QUINTNET_OUTPUT_DIR", None)
        else:
            os.environ["QUINTNET_OUTPUT_DIR"] = env_before
    vol.push("checkpoints/")
    return rc


def list_checkpoints(vol: Volume, run: Optional[str] = None) -> List[dict]:
    vol.pull("checkpoints/")
    rows = _walk(vol.path("checkpoints", run) if run else vol.path("checkpoints"))
    for r in rows:
        print(f"{r['bytes'] / 1e6:10.2f} MB  {r['mtime']}  {r['path']}")
    return rows


def download_checkpoints(vol: Volume, dst: str, run: Optional[str] = None) -> str:
    vol.pull("checkpoints/")
    src = vol.path("checkpoints", run) if run else vol.path("checkpoints")
    shutil.copytree(src, dst, dirs_exist_ok=True)
    return dst


def test_weight_shapes(model: str, tp: int, pp: int, checkpoint: Optional[str]) -> int:
    from . import check_weight_shapes
    argv = ["--model", model, "--tp", str(tp), "--pp", str(pp)] + (["--checkpoint", checkpoint] if checkpoint else [])
    return check_weight_shapes.main(argv) or 0


def merge_and_test(vol: Volume, run: str, prefix: str, dataset: Optional[str], max_samples: int, generation: bool) -> dict:
    """``merge_and_test_model``: shards → one HuggingFace-layout state dict, then perplexity (and optionally ROUGE/BLEU of
    greedy summaries) of the merged model."""
    from . import eval_checkpoint, merge_checkpoints
    vol.pull(f"checkpoints/{run}/")
    merged = vol.path("checkpoints", run, f"{prefix}_merged.pt")
    rc = merge_checkpoints.main(["--input_dir", vol.path("checkpoints", run), "--output", merged, "--prefix", prefix]) or 0
    if rc:
        return {"merge_rc": rc}
    argv = ["--checkpoint", merged, "--max_samples", str(max_samples)]
    if dataset:
        argv += ["--dataset", vol.path("datasets", dataset)]
    if generation:
        argv += ["--generation"]
    res = eval_checkpoint.main(argv)
    vol.push(f"checkpoints/{run}/")
    return {"merge_rc": 0, "merged": merged, "eval": res}


def run_tests(pytest_args: List[str]) -> int:
    """``test_modal_run.py``: run (a subset of) the test suite on this box."""
    return subprocess.call([sys.executable, "-m", "pytest", *(pytest_args or ["tests", "-q", "-m", "not gpu"])])


# -------------------------------------------------------------------------------------------------- CLI
def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--volume", default=os.environ.get("QUINTNET_VOLUME", "./qn_volume"))
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("upload-model"); p.add_argument("--src", required=True)
    p = sub.add_parser("upload-dataset"); p.add_argument("--src", required=True)
    p = sub.add_parser("prepare-dataset")
    p.add_argument("--name", default="cnn_dailymail"); p.add_argument("--tokenizer", default="gpt2")
    p.add_argument("--max-length", type=int, default=512); p.add_argument("--max-samples", type=int, default=None)
    p.add_argument("--synthetic", action="store_true")
    p = sub.add_parser("train")
    p.add_argument("--workload", choices=["vit", "gpt2"], default="gpt2"); p.add_argument("--nproc", type=int, default=8)
    p.add_argument("--run", default=time.strftime("run_%Y%m%d_%H%M%S")); p.add_argument("--config", default=None)
    p.add_argument("--port", type=int, default=29581); p.add_argument("rest", nargs=argparse.REMAINDER)
    p = sub.add_parser("list-checkpoints"); p.add_argument("--run", default=None)
    p = sub.add_parser("download-checkpoints"); p.add_argument("--dst", required=True); p.add_argument("--run", default=None)
    p = sub.add_parser("test-weight-shapes")
    p.add_argument("--model", default="gpt2_base"); p.add_argument("--tp", type=int, default=2)
    p.add_argument("--pp", type=int, default=1); p.add_argument("--checkpoint", default=None)
    p = sub.add_parser("merge-and-test")
    p.add_argument("--run", required=True); p.add_argument("--prefix", default="final_model")
    p.add_argument("--dataset", default=None); p.add_argument("--max-samples", type=int, default=256)
    p.add_argument("--generation", action="store_true")
    p = sub.add_parser("test"); p.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    vol = Volume(a.volume)
    if a.cmd == "upload-model":
        print(upload_model(vol, a.src))
    elif a.cmd == "upload-dataset":
        print(upload_dataset(vol, a.src))
    elif a.cmd == "prepare-dataset":
        print(prepare_dataset(vol, a.name, a.tokenizer, a.max_length, a.max_samples, a.synthetic))
    elif a.cmd == "train":
        return train(vol, a.workload, a.nproc, a.run, a.config, [x for x in a.rest if x != "--"], a.port)
    elif a.cmd == "list-checkpoints":
        list_checkpoints(vol, a.run)
    elif a.cmd == "download-checkpoints":
        print(download_checkpoints(vol, a.dst, a.run))
    elif a.cmd == "test-weight-shapes":
        return test_weight_shapes(a.model, a.tp, a.pp, a.checkpoint)
    elif a.cmd == "merge-and-test":
        print(json.dumps(merge_and_test(vol, a.run, a.prefix, a.dataset, a.max_samples, a.generation), default=str))
    elif a.cmd == "test":
        return run_tests([x for x in a.rest if x != "--"])
    return 0


if __name__ == "__main__":
    sys.exit(main())
