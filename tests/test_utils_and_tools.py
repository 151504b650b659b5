"""Utilities, tools and examples: datasets + collator, metrics, logging, profiling,
single-device checkpoint evaluation, and the shipped examples in CPU / gloo mode."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _WordTokenizer:
    """Whitespace tokenizer with the slice of the HF tokenizer API the collator uses."""
    eos_token, pad_token = "<eos>", None

    def __init__(self):
        self.vocab = {"<eos>": 0}

    def _ids(self, text):
        return [self.vocab.setdefault(w, len(self.vocab)) for w in text.replace("<eos>", " <eos>").split()]

    def __call__(self, texts, max_length=None, truncation=False, padding=None, return_tensors=None):
        single = isinstance(texts, str)
        rows = [self._ids(t)[:max_length] for t in ([texts] if single else texts)]
        if padding == "max_length":
            mask = [[1] * len(r) + [0] * (max_length - len(r)) for r in rows]
            rows = [r + [0] * (max_length - len(r)) for r in rows]
        else:
            mask = [[1] * len(r) for r in rows]
        if return_tensors == "pt":
            return {"input_ids": torch.tensor(rows), "attention_mask": torch.tensor(mask)}
        return {"input_ids": rows[0] if single else rows, "attention_mask": mask[0] if single else mask}


def test_summarization_dataset_and_collator(tmp_path):
    from quintnet_b200.utils.data import SummarizationCollator, SummarizationDataset
    (tmp_path / "train.csv").write_text("article,highlights\nthe cat sat on the mat,cat sat\na dog ran,dog\n")
    ds = SummarizationDataset(str(tmp_path), "train")
    assert len(ds) == 2 and ds[0] == {"article": "the cat sat on the mat", "highlights": "cat sat"}
    tok = _WordTokenizer()
    batch = SummarizationCollator(tok, max_length=16)([ds[0], ds[1]])
    assert batch["input_ids"].shape == (2, 16) and tok.pad_token == tok.eos_token
    assert (batch["labels"][batch["attention_mask"] == 0] == -100).all()        # padding is not supervised
    assert (batch["labels"][batch["attention_mask"] == 1] >= 0).all()           # the whole text is (reference behaviour)
    masked = SummarizationCollator(tok, max_length=16, mask_prompt=True)([ds[0]])
    n_prompt = len(tok("the cat sat on the mat\n\nTL;DR: ")["input_ids"])
    assert (masked["labels"][0, :n_prompt] == -100).all() and (masked["labels"][0, n_prompt:n_prompt + 2] >= 0).all()


def test_metrics_rouge_bleu():
    from quintnet_b200.utils.metrics import compute_rouge_bleu
    same = compute_rouge_bleu(["the cat sat on the mat"], ["the cat sat on the mat"])
    diff = compute_rouge_bleu(["a dog ran away quickly today"], ["the cat sat on the mat"])
    for key in ("rouge1", "rouge2", "rougeL", "bleu"):
        assert key in same and same[key] >= diff[key]
    assert same["rouge1"] > 0.99 and diff["rouge1"] < 0.2


def test_loggers(tmp_path, monkeypatch):
    from quintnet_b200.utils.logger import StepLogger, log_rank_0, setup_logger, setup_rank_logging
    sl = StepLogger(str(tmp_path / "steps.jsonl"), rank=0)
    sl.log(step=1, loss=2.5, tokens_per_s=1e5)
    sl.log(step=2, loss=2.4)
    rows = [json.loads(l) for l in (tmp_path / "steps.jsonl").read_text().splitlines()]
    assert [r["step"] for r in rows] == [1, 2] and "time" in rows[0]
    assert StepLogger(str(tmp_path / "other.jsonl"), rank=1).f is None          # rank 0 only
    monkeypatch.setenv("RANK", "3")
    out, err = sys.stdout, sys.stderr
    try:
        path = setup_rank_logging(str(tmp_path / "logs"))
        print("hello from rank 3")
        sys.stdout.flush()
    finally:
        sys.stdout, sys.stderr = out, err
    assert path.endswith("rank_3.log") and "hello from rank 3" in open(path).read()
    log_rank_0("not printed on rank 3", setup_logger("t"))


def test_profiling_and_memory():
    from quintnet_b200.utils.memory import clear_cache, get_memory_usage
    from quintnet_b200.utils.profiling import DeviceTimer, nvtx_range, profile_memory, profile_time
    sink = {}
    with profile_time("mm", sink=sink), profile_memory(name="mm_mem", sink=sink), nvtx_range("mm"):
        torch.randn(64, 64) @ torch.randn(64, 64)
    assert sink["mm"] > 0 and sink["mm_mem"] >= 0
    t = DeviceTimer()
    with t.phase("fwd"):
        torch.randn(32, 32).sum()
    with t.phase("fwd"):
        torch.randn(32, 32).sum()
    rep = t.summary()
    assert rep["fwd"] >= 0 and t._counts["fwd"] == 2
    assert isinstance(get_memory_usage(), dict)
    clear_cache()


def test_eval_checkpoint_tool(tmp_path, capsys):
    """tools/eval_checkpoint (the reference's test.py role): loss / perplexity of an HF-layout checkpoint."""
    from quintnet_b200.core.distributed_loading import save_random_hf_checkpoint
    from quintnet_b200.models.gpt2 import GPT2Config
    from quintnet_b200.tools import eval_checkpoint
    cfg = GPT2Config(vocab_size=128, n_positions=32, n_embd=64, n_layer=2, n_head=1)
    path = save_random_hf_checkpoint(str(tmp_path / "tiny.safetensors"), cfg, seed=3)
    eval_checkpoint.main(["--checkpoint", path, "--max_samples", "16"])
    out = capsys.readouterr().out
    assert "perplexity" in out
    loss = float(out.split("loss")[1].split("|")[0])
    assert 3.0 < loss < 7.0                                     # ~ln(128) = 4.85 for a random model


@pytest.mark.parametrize("example,nproc", [("examples.simple_pp", 2), ("examples.simple_tp", 2),   # simple_dp: README-line test below
                                           ("examples.full_3d", 8)])
def test_examples_run_on_cpu(example, nproc, tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(29300 + abs(hash(example)) % 500), "-m", example, "--cpu", "--synthetic",
           "--epochs", "1", "--samples", "128"]
    env = dict(os.environ, QUINTNET_OUTPUT_DIR=str(tmp_path))        # model_checkpoint.pt goes there, not into the repo
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert "Train Loss:" in out.stdout and "Val Loss:" in out.stdout          # the reference's de-facto print interface


@pytest.mark.parametrize("virtual", [1, 2])
def test_gpt2_finetune_example_cpu(tmp_path, virtual):
    """examples.gpt2_finetune end to end on 2 CPU ranks (pp=2): sharded random init -> fit -> val -> shard files;
    ``virtual=2`` runs the same through the interleaved schedule with two model chunks per rank."""
    import yaml
    cfg = {"batch_size": 8, "grad_acc_steps": 2, "num_epochs": 1, "learning_rate": 1e-3, "max_grad_norm": 1.0,
           "num_workers": 0, "max_seq_length": 16, "max_target_length": 8, "max_samples": 32, "max_val_samples": 8,
           "output_dir": str(tmp_path / "ckpt"), "compute_generation_metrics": False, "vocab_size": 96,
           "mesh_dim": [1, 1, 2], "mesh_name": ["dp", "tp", "pp"], "strategy_name": "3d", "schedule": "1f1b",
           "device_type": "cpu", "dtype": "float32", "log_jsonl": str(tmp_path / "train.jsonl"), "eval_interval": 2,
           "model_config": {"vocab_size": 96, "n_positions": 16, "n_embd": 32, "n_layer": 4, "n_head": 2},
           "virtual_pipeline_stages": virtual}
    path = tmp_path / "tiny_gpt2.yaml"
    path.write_text(yaml.safe_dump(cfg))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29877 + virtual), "-m", "examples.gpt2_finetune", "--config", str(path), "--cpu", "--synthetic"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert "Train Loss:" in out.stdout and "Val Loss:" in out.stdout
    files = sorted(os.listdir(tmp_path / "ckpt"))
    for stage in range(2 * virtual):
        assert f"final_model_pp{stage}_tp0.pt" in files, files
    # the merge CLI turns the shards into one HF-layout file (any number of stages, incl. virtual ones)
    merged = tmp_path / "merged.pt"
    cli = subprocess.run([sys.executable, "-m", "quintnet_b200.tools.merge_checkpoints", "--input_dir",
                          str(tmp_path / "ckpt"), "--output", str(merged), "--prefix", "final_model"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert cli.returncode == 0 and merged.exists(), cli.stderr[-1000:]
    hf = torch.load(merged, weights_only=False)["model_state_dict"]
    assert "transformer.h.3.mlp.c_proj.weight" in hf and "lm_head.weight" in hf
    rows = [json.loads(l) for l in (tmp_path / "train.jsonl").read_text().splitlines()]
    steps = [r for r in rows if "step" in r and "loss" in r]
    assert len(steps) == 4 and all(r["tokens_per_s"] > 0 and r["step_ms"] > 0 for r in steps)    # 32 samples / batch 8
    assert all(r["model_tflops"] > 0 and r["max_mem_gb"] is None for r in steps)                 # CPU run: no device memory
    assert [r["step"] for r in rows if "step" in r and "val_loss" in r] == [2, 4]                # eval_interval = 2
    assert any("val_loss" in r for r in rows)                                                    # epoch record (rank 0)


def test_single_device_train_then_verify(tmp_path):
    """examples.train_on_single_gpu writes model_checkpoint.pt; examples.verify_model reloads and re-measures it."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-m", "examples.train_on_single_gpu", "--cpu", "--synthetic", "--epochs", "1",
                          "--samples", "256"], capture_output=True, text=True, timeout=600, cwd=tmp_path, env=env)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert (tmp_path / "model_checkpoint.pt").exists()
    out = subprocess.run([sys.executable, "-m", "examples.verify_model", "--synthetic", "--checkpoint",
                          str(tmp_path / "model_checkpoint.pt")], capture_output=True, text=True, timeout=600,
                         cwd=tmp_path, env=env)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert "0 missing, 0 unexpected" in out.stdout and "accuracy:" in out.stdout


def test_launch_tool_parses_epoch_metrics(tmp_path):
    """tools.launch (the reference's Modal launchers' job): run an example under torchrun, parse the printed epoch
    lines into one JSON summary."""
    out = subprocess.run([sys.executable, "-m", "quintnet_b200.tools.launch", "--nproc", "2", "--port", "29691", "-m",
                          "examples.simple_dp", "--", "--cpu", "--synthetic", "--epochs", "2", "--samples", "128"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=dict(os.environ, QUINTNET_OUTPUT_DIR=str(tmp_path)))
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    summary = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert summary["exit_code"] == 0 and len(summary["epochs"]) == 2
    assert {"train_loss", "train_acc", "val_loss", "val_acc"} <= set(summary["epochs"][0])


def test_custom_dataset_from_disk(tmp_path):
    """CustomDataset over an on-disk HuggingFace dataset with PIL images (the MNIST path of the reference)."""
    datasets = pytest.importorskip("datasets")
    import numpy as np
    from PIL import Image
    from quintnet_b200.utils.data import CustomDataset, mnist_transform
    rng = np.random.default_rng(0)
    imgs = [Image.fromarray(rng.integers(0, 255, (28, 28), dtype=np.uint8)) for _ in range(6)]
    ds = datasets.DatasetDict({"train": datasets.Dataset.from_dict({"image": imgs[:4], "label": [0, 1, 2, 3]}),
                               "test": datasets.Dataset.from_dict({"image": imgs[4:], "label": [4, 5]})})
    ds.save_to_disk(str(tmp_path / "mnist"))
    train = CustomDataset(str(tmp_path / "mnist"), "train", mnist_transform)
    test = CustomDataset(str(tmp_path / "mnist"), "test", mnist_transform)
    assert len(train) == 4 and len(test) == 2
    s = train[2]
    assert s["image"].shape == (1, 28, 28) and s["image"].dtype == torch.float32 and int(s["label"]) == 2
    assert abs(float(s["image"].mean()) - (np.array(imgs[2], dtype="float32").mean() / 255 - 0.1307) / 0.3081) < 1e-4


def test_greedy_generation_and_metrics():
    """generate_summary / evaluate_generation (greedy, full re-forward per token as in the reference) on a tiny
    random GPT-2 with a toy tokenizer."""
    from quintnet_b200.models.gpt2 import GPT2Config, GPT2Model
    from quintnet_b200.utils.metrics import evaluate_generation, generate_summary

    class Tok(_WordTokenizer):
        eos_token_id = 0

        def __call__(self, texts, **kw):
            kw.pop("return_tensors", None)
            return super().__call__([texts] if isinstance(texts, str) else texts, return_tensors="pt", **kw)

        def decode(self, ids, skip_special_tokens=True):
            inv = {v: k for k, v in self.vocab.items()}
            return " ".join(inv.get(int(i), "?") for i in ids if not (skip_special_tokens and int(i) == 0))

    tok = Tok()
    for w in "the cat sat on the mat a dog ran TL;DR:".split():
        tok._ids(w)
    model = GPT2Model(GPT2Config(vocab_size=32, n_positions=32, n_embd=32, n_layer=1, n_head=2), seed=0).eval()
    text = generate_summary(model, tok, "the cat sat", max_new_tokens=5, device=torch.device("cpu"))
    assert isinstance(text, str) and len(text.split()) <= 5
    scores = evaluate_generation(model, tok, [{"article": "the cat sat on the mat", "highlights": "cat sat"}],
                                 torch.device("cpu"), num_samples=1, max_new_tokens=4)
    assert {"rouge1", "rouge2", "rougeL", "bleu"} <= set(scores)


def test_full_3d_staged_build_from_safetensors(tmp_path):
    """examples.full_3d --staged --checkpoint: every rank loads only its shard of an HF-layout safetensors file
    (tp2 x pp2 here) and reports its stage."""
    import yaml
    from quintnet_b200.core.distributed_loading import save_random_hf_checkpoint
    from quintnet_b200.models.gpt2 import GPT2Config
    mc = {"vocab_size": 96, "n_positions": 16, "n_embd": 32, "n_layer": 4, "n_head": 4}
    ck = save_random_hf_checkpoint(str(tmp_path / "tiny.safetensors"), GPT2Config(**mc), seed=5, prefix="transformer.")
    cfg = {"mesh_dim": [1, 2, 2], "mesh_name": ["dp", "tp", "pp"], "device_type": "cpu", "dtype": "float32",
           "strategy_name": "3d", "model_config": mc, "max_seq_length": 16, "batch_size": 4, "grad_acc_steps": 2}
    (tmp_path / "g.yaml").write_text(yaml.safe_dump(cfg))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
           "--master-port", "29733", "-m", "examples.full_3d", "--staged", "--config", str(tmp_path / "g.yaml"), "--cpu",
           "--checkpoint", ck]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert out.stdout.count("stage built:") == 4


def test_check_weight_shapes_tool(tmp_path, capsys):
    """tools.check_weight_shapes (the reference's test_weight_shapes.py role) on random weights and on a safetensors file."""
    from quintnet_b200.core.distributed_loading import save_random_hf_checkpoint
    from quintnet_b200.models.gpt2 import GPT2Config
    from quintnet_b200.tools import check_weight_shapes as cws
    cfg = GPT2Config(vocab_size=96, n_positions=16, n_embd=32, n_layer=4, n_head=4)
    from quintnet_b200.core.distributed_loading import hf_shapes, random_hf_tensor
    shapes = hf_shapes(cfg)
    n = cws.check(cfg, lambda name: random_hf_tensor(name, shapes[name], cfg, 0, torch.device("cpu")), tp=2, pp=2, verbose=False)
    assert n > 40
    cws.main(["--model", "gpt2_base", "--tp", "4", "--pp", "2"])
    assert "shard shapes match" in capsys.readouterr().out


def test_reference_readme_launch_lines_work(tmp_path):
    """The reference README's launch line, verbatim apart from the CPU flags: ``-m QuintNet.examples.simple_dp`` with
    ``--config QuintNet/examples/dp_config.yaml``."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29612", "-m", "QuintNet.examples.simple_dp", "--config", "QuintNet/examples/dp_config.yaml",
           "--cpu", "--synthetic", "--epochs", "1", "--samples", "64"]
    env = dict(os.environ, QUINTNET_OUTPUT_DIR=str(tmp_path))      # the relative config path needs cwd = repo root;
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)   # the checkpoint does not
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert "Train Loss:" in out.stdout
    assert (tmp_path / "model_checkpoint.pt").exists()


def test_cluster_workflow_on_a_volume_directory(tmp_path):
    """tools.cluster — the reference's Modal workflow (upload_model / prepare_dataset / run_gpt2_training /
    list_checkpoints / download_checkpoints / test_weight_shapes / merge_and_test_model / run_test) against a plain
    volume directory: stage -> train (torchrun, 2 CPU ranks, pp=2) -> list -> merge + evaluate -> download."""
    import yaml
    from quintnet_b200.core.distributed_loading import save_random_hf_checkpoint
    from quintnet_b200.models.gpt2 import GPT2Config
    vol = tmp_path / "vol"
    mc = {"vocab_size": 96, "n_positions": 16, "n_embd": 32, "n_layer": 2, "n_head": 2}
    src = tmp_path / "tiny.safetensors"
    save_random_hf_checkpoint(str(src), GPT2Config.from_dict(mc), seed=0)
    cfg = {"batch_size": 8, "grad_acc_steps": 2, "num_epochs": 1, "learning_rate": 1e-3, "max_grad_norm": 1.0,
           "num_workers": 0, "max_seq_length": 16, "max_target_length": 8, "max_samples": 16, "max_val_samples": 8,
           "compute_generation_metrics": False, "vocab_size": 96, "mesh_dim": [1, 1, 2], "mesh_name": ["dp", "tp", "pp"],
           "strategy_name": "3d", "schedule": "1f1b", "device_type": "cpu", "dtype": "float32", "model_config": mc}
    cpath = tmp_path / "tiny.yaml"
    cpath.write_text(yaml.safe_dump(cfg))
    base = [sys.executable, "-m", "quintnet_b200.tools.cluster", "--volume", str(vol)]
    run = lambda *a, **kw: subprocess.run(base + list(a), capture_output=True, text=True, timeout=600, cwd=ROOT, **kw)
    r = run("upload-model", "--src", str(src))
    assert r.returncode == 0 and (vol / "models" / "tiny.safetensors").exists(), r.stderr[-800:]
    r = run("prepare-dataset", "--name", "toy", "--synthetic", "--max-length", "16", "--max-samples", "32")
    assert r.returncode == 0 and (vol / "datasets" / "toy_tok16.pt").exists(), r.stderr[-800:]
    r = run("train", "--workload", "gpt2", "--nproc", "2", "--run", "exp1", "--config", str(cpath), "--port", "29717",
            "--", "--cpu", "--synthetic")
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    summary = json.loads(r.stdout.strip().splitlines()[-1])
    assert summary["exit_code"] == 0 and summary["epochs"] and "train_loss" in summary["epochs"][0]
    r = run("list-checkpoints", "--run", "exp1")
    assert "final_model_pp0_tp0.pt" in r.stdout and "final_model_pp1_tp0.pt" in r.stdout, r.stdout
    r = run("merge-and-test", "--run", "exp1", "--prefix", "final_model", "--max-samples", "8")
    assert r.returncode == 0 and (vol / "checkpoints" / "exp1" / "final_model_merged.pt").exists(), (r.stdout[-800:], r.stderr[-800:])
    assert "perplexity" in r.stdout
    r = run("download-checkpoints", "--run", "exp1", "--dst", str(tmp_path / "dl"))
    assert r.returncode == 0 and (tmp_path / "dl" / "final_model_merged.pt").exists()
    r = run("test-weight-shapes", "--model", "gpt2_base", "--tp", "2", "--pp", "2")
    assert r.returncode == 0, r.stderr[-800:]


def test_symmetric_buffer_lookup_bisects_the_registry():
    """ops.symm.lookup maps a data pointer to (buffer, byte offset) by bisecting the sorted base addresses."""
    import quintnet_b200.ops.symm as symm

    class Buf:
        def __init__(self, base, n):
            self.base, self.nbytes = base, n

    class Ptr:
        is_cuda = True

        def __init__(self, p):
            self.p = p

        def data_ptr(self):
            return self.p

    saved, symm._registry[:] = list(symm._registry), [Buf(5000, 100), Buf(1000, 50), Buf(3000, 10)]
    symm._sorted_bases = None
    try:
        for p, want in [(1000, (1000, 0)), (1049, (1000, 49)), (1050, None), (999, None), (3005, (3000, 5)),
                        (5099, (5000, 99)), (5100, None)]:
            r = symm.lookup(Ptr(p))
            assert (None if r is None else (r[0].base, r[1])) == want, p
        symm._registry.append(Buf(7000, 10))                 # a later allocation is picked up
        assert symm.lookup(Ptr(7003))[1] == 3
    finally:
        symm._registry[:] = saved
        symm._sorted_bases = None
