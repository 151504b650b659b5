"""Public-API parity, config handling, checkpoint shard -> merge -> HF round trip, resume."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist

from tests._dist import run_distributed


def test_reference_import_surface():
    import QuintNet
    from QuintNet import (init_process_groups, get_strategy, Trainer, setup_distributed, cleanup_distributed,  # noqa
                          DataParallel, TensorParallel, PipelineParallelWrapper, PipelineTrainer, PipelineDataLoader,
                          BaseStrategy, DataParallelStrategy, TensorParallelStrategy, PipelineParallelStrategy,
                          DataTensorParallelStrategy, DataPipelineParallelStrategy, TensorPipelineParallelStrategy,
                          Hybrid3DStrategy)
    from QuintNet.core import (MeshGenerator, ProcessGroupManager, load_config, All_Reduce, All_Gather,  # noqa
                               ReduceScatter, Send, Recv, pipeline_communicate, bidirectional_pipeline_communicate,
                               load_gpt2_distributed)
    from QuintNet.parallelism import ColumnParallelLinear, RowParallelLinear  # noqa
    from QuintNet.parallelism.data_parallel import (GradientBucket, BucketManager, GradientReducer,  # noqa
                                                    ParameterBroadcaster, create_local_ddp, create_distributed_ddp,
                                                    DistributedConfig, BucketConfig, ReductionStrategy)
    from QuintNet.parallelism.tensor_parallel import VocabParallelEmbedding  # noqa
    from QuintNet.utils.GPT2 import GPT2Config, GPT2Embedding, GPT2Attention, GPT2MLP, GPT2Block, GPT2Stage  # noqa
    from QuintNet.utils.model import Model  # noqa
    from QuintNet.GPT2_Trainer import GPT2Trainer  # noqa
    assert QuintNet.__version__


def test_get_strategy_names():
    from quintnet_b200 import get_strategy, init_process_groups
    pg = init_process_groups("cpu", [1, 1, 1], ["dp", "tp", "pp"])
    for name in ["dp", "tp", "pp", "dp_tp", "dp_pp", "tp_pp", "3d", "3D", "Dp_Tp"]:
        assert get_strategy(name, pg, {"device_type": "cpu"}) is not None
    with pytest.raises(ValueError):
        get_strategy("zero", pg, {})


def test_config_load_merge(tmp_path):
    from quintnet_b200.core.config import load_config, merge_configs, axis_size, parallelism_view
    p = tmp_path / "c.yaml"
    p.write_text("learning_rate: 1e-4\nmesh_dim: [2, 4]\nmesh_name: ['dp', 'pp']\nmodel_config:\n  n_embd: 64\n")
    cfg = load_config(str(p))
    assert cfg["learning_rate"] == pytest.approx(1e-4)
    assert axis_size(cfg, "pp") == 4 and axis_size(cfg, "tp") == 1
    merged = merge_configs(cfg, {"model_config": {"n_layer": 2}, "learning_rate": 3e-4})
    assert merged["model_config"] == {"n_embd": 64, "n_layer": 2} and merged["learning_rate"] == 3e-4
    assert parallelism_view(cfg).axis("dp") == 2
    with pytest.raises(FileNotFoundError):
        load_config(str(tmp_path / "missing.yaml"))


CFG = dict(vocab_size=96, n_positions=16, n_embd=32, n_layer=4, n_head=4, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)


def _train_cfg(out_dir, zero=0):
    return {"batch_size": 8, "learning_rate": 1e-2, "num_epochs": 1, "grad_acc_steps": 2, "max_grad_norm": 1.0,
            "model_config": dict(CFG), "device_type": "cpu", "mesh_dim": [1, 2, 2], "mesh_name": ["dp", "tp", "pp"],
            "strategy_name": "3d", "schedule": "1f1b", "max_seq_length": 16, "zero_stage": zero, "seed": 1,
            "output_dir": out_dir}


def _save_merge(rank, world, out_dir):
    from quintnet_b200 import GPT2Trainer, get_strategy, init_process_groups
    from quintnet_b200.models.gpt2 import GPT2Config
    from quintnet_b200.utils.data import SyntheticTokens, make_loader
    cfg = _train_cfg(out_dir)
    pg = init_process_groups("cpu", cfg["mesh_dim"], cfg["mesh_name"])
    model = get_strategy("3d", pg, cfg, None, is_staged=True).apply(None)
    loader = make_loader(SyntheticTokens(32, 16, 96, seed=2, pad_id=0), 4, pg, shuffle=False)
    tr = GPT2Trainer(model, loader, None, cfg, pg, GPT2Config(**CFG))
    for _ in range(2):
        tr.train_step()
    tr._save_checkpoint("final_model.pt")
    dist.barrier()
    # resume into a fresh trainer: parameters and optimizer state come back bit-exactly
    model2 = get_strategy("3d", pg, dict(cfg, seed=7), None, is_staged=True).apply(None)
    tr2 = GPT2Trainer(model2, loader, None, cfg, pg, GPT2Config(**CFG))
    tr2.resume(os.path.join(out_dir, "final_model.pt"))
    for (n1, p1), (n2, p2) in zip(tr.gpt2_stage.named_parameters(), tr2.gpt2_stage.named_parameters()):
        assert torch.equal(p1, p2), n1
    assert tr2.global_step == 2
    from quintnet_b200.parallel.pipeline_parallel import PipelineDataLoader
    tr.micro_loader = PipelineDataLoader(loader, 2, "clm")       # same data for both continuations
    tr2.micro_loader = PipelineDataLoader(loader, 2, "clm")
    l1, l2 = tr.train_step(), tr2.train_step()
    if l1[0] is not None:
        assert abs(l1[0] - l2[0]) < 1e-5, "resumed run continues identically"


def test_shards_merge_to_hf_and_resume():
    transformers = pytest.importorskip("transformers")
    from quintnet_b200.tools.merge_checkpoints import merge_checkpoints
    from quintnet_b200.models.gpt2 import GPT2Config, GPT2Model
    with tempfile.TemporaryDirectory() as d:
        run_distributed(_save_merge, 4, d)
        files = sorted(os.listdir(d))
        assert files == [f"final_model_pp{p}_tp{t}.pt" for p in range(2) for t in range(2)], files
        shard = torch.load(os.path.join(d, files[0]), weights_only=False)
        assert {"model_state_dict", "optimizer_state_dict", "config", "parallelism_info"} <= set(shard)
        assert "blocks.0.attn.c_attn.proj.weight" in shard["model_state_dict"]
        out = os.path.join(d, "merged.pt")
        hf_sd = merge_checkpoints(d, out, "final_model")
        assert hf_sd["transformer.h.3.attn.c_attn.weight"].shape == (32, 96)      # Conv1D [in, out]
        hf = transformers.GPT2LMHeadModel(transformers.GPT2Config(
            vocab_size=96, n_positions=16, n_embd=32, n_layer=4, n_head=4, activation_function="gelu",
            resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)).eval()
        merged = torch.load(out, weights_only=False)
        missing, unexpected = hf.load_state_dict(merged["model_state_dict"], strict=False)
        assert not unexpected and all("attn.bias" in m or "masked_bias" in m for m in missing), (missing, unexpected)
        # the merged model computes the same function as an un-sharded model with the merged weights
        ids = torch.randint(0, 96, (2, 16))
        ours = GPT2Model(GPT2Config(**CFG), seed=1).eval()
        sd = ours.state_dict()
        with torch.no_grad():
            logits_hf = hf(ids).logits
        # rebuild our single-device model from the HF tensors via the sharded loader path
        from quintnet_b200.core.distributed_loading import shard_hf_tensors
        full = {k.replace("transformer.", ""): v for k, v in merged["model_state_dict"].items()}
        local = shard_hf_tensors(lambda n: full[n], GPT2Config(**CFG), 0, 1, 0, 1, torch.device("cpu"))
        from quintnet_b200.models.gpt2 import GPT2Stage
        m = GPT2Stage.from_sharded_state_dict(local, GPT2Config(**CFG), 0, 1, 0, 1, None, None, torch.device("cpu")).eval()
        with torch.no_grad():
            assert torch.allclose(m(ids), logits_hf, atol=2e-4)


def _vit_trainer(rank, world, strategy, mesh_dim, mesh_name):
    from quintnet_b200 import Trainer, get_strategy, init_process_groups
    from quintnet_b200.models import Model
    from quintnet_b200.utils.data import SyntheticMNIST, make_loader
    cfg = {"batch_size": 16, "learning_rate": 3e-3, "num_epochs": 2, "grad_acc_steps": 2, "max_grad_norm": 1.0,
           "img_size": 28, "patch_size": 4, "hidden_dim": 32, "in_channels": 1, "n_heads": 4, "depth": 4,
           "device_type": "cpu", "mesh_dim": mesh_dim, "mesh_name": mesh_name, "strategy_name": strategy,
           "schedule": "1f1b", "output_dir": tempfile.mkdtemp(), "tp_method": "paired"}
    pg = init_process_groups("cpu", mesh_dim, mesh_name)
    dp = pg.axis_size("dp")
    mb = 16 // (2 * dp)
    tl = make_loader(SyntheticMNIST(256, seed=0), mb, pg)
    vl = make_loader(SyntheticMNIST(64, seed=1), mb, pg, shuffle=False)
    torch.manual_seed(0)
    model = Model(hidden_dim=32, depth=4, n_heads=4)
    model = get_strategy(strategy, pg, cfg).apply(model)
    tr = Trainer(model, tl, vl, cfg, pg)
    l0, _ = tr._train_epoch(0)
    l1, _ = tr._train_epoch(1)
    if l0 is not None:
        assert l1 < l0, (l0, l1)
    v = tr._validate_epoch(0)
    assert v is not None


@pytest.mark.parametrize("strategy,mesh_dim,mesh_name,world", [
    ("dp", [2], ["dp"], 2), ("tp", [2], ["tp"], 2), ("pp", [2], ["pp"], 2),
    ("dp_tp", [2, 2], ["dp", "tp"], 4), ("dp_pp", [2, 2], ["dp", "pp"], 4), ("tp_pp", [2, 2], ["tp", "pp"], 4),
    ("3d", [2, 2, 2], ["dp", "tp", "pp"], 8)])
def test_vit_trainer_all_strategies(strategy, mesh_dim, mesh_name, world):
    run_distributed(_vit_trainer, world, strategy, mesh_dim, mesh_name, timeout=400)


def test_bench_contract_cpu():
    """bench.py prints ONE JSON line with the keys the driver reads (CPU plumbing mode, tiny model)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--cpu", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "impl"):
        assert key in rec, key
    assert rec["n_gpus"] == 1 and rec["steps"] == 1 and rec["higher_is_better"] is True and rec["scaling"] == "strong"
    assert {"model", "global_batch", "seq_len", "parallelism", "l2"} <= set(rec["config"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(rec["e2e"])
    assert rec["value"] > 0 and rec["e2e"]["h2d_bytes_per_step"] > 0


def _zero1_resume(rank, world, out_dir):
    """ZeRO-1 under dp2 x pp2: every dp rank writes its optimizer shard; a fresh trainer resumes bit-exactly."""
    from quintnet_b200 import GPT2Trainer, get_strategy, init_process_groups
    from quintnet_b200.models.gpt2 import GPT2Config
    from quintnet_b200.parallel.pipeline_parallel import PipelineDataLoader
    from quintnet_b200.utils.data import SyntheticTokens, make_loader
    cfg = dict(_train_cfg(out_dir, zero=1), mesh_dim=[2, 1, 2])
    pg = init_process_groups("cpu", cfg["mesh_dim"], cfg["mesh_name"])
    model = get_strategy("3d", pg, cfg, None, is_staged=True).apply(None)
    loader = make_loader(SyntheticTokens(32, 16, 96, seed=2, pad_id=0, min_fill=1.0), 2, pg, shuffle=False)
    tr = GPT2Trainer(model, loader, None, cfg, pg, GPT2Config(**CFG))
    for _ in range(2):
        tr.train_step()
    tr._save_checkpoint("z1.pt")
    dist.barrier()
    if dist.get_rank() == 0:
        names = sorted(os.listdir(out_dir))
        assert sum(n.endswith(".opt.pt") for n in names) == 4, names          # one optimizer shard per (pp, dp) rank
    model2 = get_strategy("3d", pg, dict(cfg, seed=7), None, is_staged=True).apply(None)
    tr2 = GPT2Trainer(model2, loader, None, cfg, pg, GPT2Config(**CFG))
    tr2.resume(os.path.join(out_dir, "z1.pt"))
    for (n1, p1), (n2, p2) in zip(tr.gpt2_stage.named_parameters(), tr2.gpt2_stage.named_parameters()):
        assert torch.equal(p1, p2), n1
    tr.micro_loader = PipelineDataLoader(loader, 2, "clm")
    tr2.micro_loader = PipelineDataLoader(loader, 2, "clm")
    l1, l2 = tr.train_step(), tr2.train_step()
    if l1[0] is not None:
        assert abs(l1[0] - l2[0]) < 1e-5
    for (n1, p1), (n2, p2) in zip(tr.gpt2_stage.named_parameters(), tr2.gpt2_stage.named_parameters()):
        assert torch.allclose(p1, p2, atol=1e-6), n1                           # optimizer state came back too


def test_zero1_checkpoint_resume():
    with tempfile.TemporaryDirectory() as d:
        run_distributed(_zero1_resume, 4, d)


def test_bench_reference_arm_reports_unavailable_without_gpu():
    """`bench.py --impl reference` must never crash the driver: where the reference cannot run (no CUDA here) it
    prints {"impl": "reference", "unavailable": ...} and exits 0."""
    import json
    import subprocess
    import sys
    if torch.cuda.is_available():
        pytest.skip("GPU present: the reference arm actually runs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-1500:]
    rec = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert rec["impl"] == "reference" and "unavailable" in rec


def test_bench_parity_mode_cpu(tmp_path):
    """`bench.py --parity-steps` + `--check-parity` (the plumbing `tests/test_parity_gpu.py` relies on): the same seeded
    global batch gives the same loss curve on one rank and on pp2 (the loss lives on the last stage; gloo)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    common = ["--cpu", "--parity-steps", "2", "--steps", "1", "--warmup", "1"]
    one = str(tmp_path / "n1.json")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), *common, "--parity-out", one],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-1500:]
    # pp2 twice: next-activation receive pre-posted before the backward (QN_PP_PREPOST_FWD=1) and posted after it (=0,
    # the order used with sequence-parallel TP) — both orders of the 1F1B steady state must give the same curve
    for i, (mesh, prepost) in enumerate((("1,1,2", "1"), ("1,1,2", "0"))):
        two = str(tmp_path / f"n2_{i}.json")
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                              "--master-addr", "127.0.0.1", "--master-port", str(29921 + i), os.path.join(root, "bench.py"),
                              *common, "--gpus", "2", "--mesh", mesh, "--parity-out", two],
                             capture_output=True, text=True, timeout=600, cwd=root,
                             env=dict(os.environ, QN_PP_PREPOST_FWD=prepost))
        assert out.returncode == 0, out.stderr[-1500:]
        chk = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--check-parity", one, two, "--parity-tol", "1e-3"],
                             capture_output=True, text=True, timeout=120, cwd=root)
        assert chk.returncode == 0, chk.stdout + chk.stderr[-500:]
        rec = json.loads(chk.stdout.strip().splitlines()[-1])
        assert rec["parity"] == "ok" and rec["max_abs_diff"] < 1e-3
    # a curve that differs is reported as a failure (non-zero exit)
    bad = json.load(open(one))
    bad["losses"] = [x + 1.0 for x in bad["losses"]]
    json.dump(bad, open(str(tmp_path / "bad.json"), "w"))
    chk = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--check-parity", one, str(tmp_path / "bad.json")],
                         capture_output=True, text=True, timeout=120, cwd=root)
    assert chk.returncode != 0


def test_bench_vit_arm_cpu():
    """`bench.py --workload vit` (BASELINE config #2's harness) prints the contract line through Trainer's step loop."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "vit", "--cpu", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-1500:]
    rec = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert rec["value"] > 0 and rec["images_per_s"] > 0 and rec["config"]["seq_len"] == 50 and rec["steps"] == 2


_REFERENCE_MODULES = """GPT2_Trainer coordinators coordinators.data_parallel_coordinator coordinators.dp_pp_coordinator
coordinators.dp_tp_coordinator coordinators.hybrid_3d_coordinator coordinators.main_coordinator
coordinators.pipeline_parallel_coordinator coordinators.tensor_parallel_coordinator coordinators.tp_pp_coordinator core
core.communication core.config core.distributed core.distributed_loading core.mesh core.process_groups merge_checkpoints
optimizers optimizers.distributed_adamw optimizers.zero parallelism parallelism.data_parallel
parallelism.data_parallel.backends parallelism.data_parallel.backends.base parallelism.data_parallel.backends.local_backend
parallelism.data_parallel.backends.nccl parallelism.data_parallel.backends.torch_backend parallelism.data_parallel.components
parallelism.data_parallel.components.bucket parallelism.data_parallel.components.bucket_manager
parallelism.data_parallel.components.gradient_reducer parallelism.data_parallel.components.parameter_broadcaster
parallelism.data_parallel.core parallelism.data_parallel.core.config parallelism.data_parallel.core.ddp
parallelism.data_parallel.utils parallelism.data_parallel.utils.factory parallelism.pipeline_parallel
parallelism.pipeline_parallel.dataloader parallelism.pipeline_parallel.schedule parallelism.pipeline_parallel.trainer
parallelism.pipeline_parallel.utils parallelism.pipeline_parallel.wrapper parallelism.tensor_parallel
parallelism.tensor_parallel.layers parallelism.tensor_parallel.model_wrapper parallelism.tensor_parallel.utils strategy
strategy.base_strategy strategy.data_parallel_strategy strategy.dp_pp_strategy strategy.dp_tp_strategy
strategy.hybrid_3d_strategy strategy.pipeline_parallel_strategy strategy.tensor_parallel_strategy strategy.tp_pp_strategy
trainer utils utils.Dataloader utils.GPT2 utils.GPT2.gpt2_attention utils.GPT2.gpt2_block utils.GPT2.gpt2_config
utils.GPT2.gpt2_embeddings utils.GPT2.gpt2_mlp utils.GPT2.gpt2_model utils.GPT2.gpt2_stage utils.checkpoint utils.logger
utils.logging utils.memory utils.metrics utils.model utils.profiling utils.utils""".split()


def test_every_reference_module_path_imports():
    """A user of the reference can keep EVERY import line: each module path of its source tree resolves under
    ``QuintNet.`` to the implementing module (the same objects, not copies)."""
    import importlib
    import QuintNet
    import quintnet_b200
    for m in _REFERENCE_MODULES:
        importlib.import_module("QuintNet." + m)
    from QuintNet.parallelism.tensor_parallel.layers import ColumnParallelLinear as a
    from quintnet_b200.parallel.tensor_parallel.layers import ColumnParallelLinear as b
    assert a is b
    from QuintNet.utils.utils import load_checkpoint, save_checkpoint          # noqa: F401
    from QuintNet.strategy.hybrid_3d_strategy import Hybrid3DStrategy          # noqa: F401
    from QuintNet.coordinators.hybrid_3d_coordinator import Hybrid3DCoordinator  # noqa: F401
    from QuintNet.GPT2_Trainer import GPT2Trainer
    assert GPT2Trainer is quintnet_b200.GPT2Trainer and QuintNet.core.mesh is quintnet_b200.core.mesh
    if os.path.isdir("/root/reference"):                 # the list above is complete w.r.t. the mounted reference tree
        have = set()
        for root, _, files in os.walk("/root/reference"):
            rel = os.path.relpath(root, "/root/reference")
            if rel.split(os.sep)[0] in ("tests", "examples", ".git"):
                continue
            for f in files:
                if f.endswith(".py") and f not in ("setup.py", "test.py", "test_weight_shapes.py") and "modal_run" not in f:
                    mod = os.path.join(rel, f[:-3]).replace(os.sep, ".").lstrip(".")
                    mod = mod[:-9] if mod.endswith(".__init__") else mod
                    if mod and mod != "__init__":
                        have.add(mod.lstrip("."))
        assert have <= set(_REFERENCE_MODULES), sorted(have - set(_REFERENCE_MODULES))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted")
def test_every_public_symbol_of_the_reference_exists():
    """Symbol-level parity, derived from the reference sources with ``ast`` only (nothing of the reference is
    imported or executed): every public top-level class / function of every module — and every name its
    ``__init__`` files re-export — exists on the module that ``QuintNet.<same path>`` resolves to."""
    import ast
    import importlib
    missing = []
    for root, _, files in os.walk("/root/reference"):
        rel = os.path.relpath(root, "/root/reference")
        if rel.split(os.sep)[0] in ("tests", "examples", ".git"):
            continue
        for f in files:
            if not f.endswith(".py") or f in ("setup.py", "test.py", "test_weight_shapes.py") or "modal_run" in f:
                continue
            mod = os.path.join(rel, f[:-3]).replace(os.sep, ".").lstrip(".")
            mod = mod[:-9] if mod.endswith(".__init__") else mod
            mod = "" if mod in ("__init__", "") else mod
            tree = ast.parse(open(os.path.join(root, f)).read())
            names = [n.name for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and not n.name.startswith("_")]
            if f == "__init__.py":
                for n in tree.body:
                    if isinstance(n, ast.ImportFrom):
                        names += [a.asname or a.name for a in n.names if a.name != "*"]
            target = importlib.import_module("QuintNet" + ("." + mod if mod else ""))
            missing += [(mod, nm) for nm in names
                        if not hasattr(target, nm) and nm not in ("Dict", "Any", "Optional", "List", "Tuple", "Union")]
    assert not missing, missing


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted")
def test_every_public_method_of_the_reference_classes_exists():
    """Method-level parity (ast only): every public method of every public reference class exists on our class."""
    import ast
    import importlib
    missing = []
    for root, _, files in os.walk("/root/reference"):
        rel = os.path.relpath(root, "/root/reference")
        if rel.split(os.sep)[0] in ("tests", "examples", ".git"):
            continue
        for f in files:
            if not f.endswith(".py") or f in ("setup.py", "test.py", "test_weight_shapes.py") or "modal_run" in f:
                continue
            mod = os.path.join(rel, f[:-3]).replace(os.sep, ".").lstrip(".")
            mod = mod[:-9] if mod.endswith(".__init__") else mod
            mod = "" if mod in ("__init__", "") else mod
            target = importlib.import_module("QuintNet" + ("." + mod if mod else ""))
            for n in ast.parse(open(os.path.join(root, f)).read()).body:
                if isinstance(n, ast.ClassDef) and not n.name.startswith("_"):
                    cls = getattr(target, n.name)
                    missing += [(mod, n.name, b.name) for b in n.body
                                if isinstance(b, ast.FunctionDef) and not b.name.startswith("_") and not hasattr(cls, b.name)]
    assert not missing, missing


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted")
def test_every_reference_signature_is_accepted():
    """Signature-level parity (ast vs inspect): for every public function, method and constructor of the reference, our
    callable accepts all of its parameter names, in the same positional order (extra parameters of ours come after
    or are keyword-only in practice)."""
    import ast
    import importlib
    import inspect
    issues = []

    def check(label, fnode, obj):
        try:
            sig = inspect.signature(obj)
        except (TypeError, ValueError):
            return
        names = [x.arg for x in fnode.args.posonlyargs + fnode.args.args]
        kwonly = [x.arg for x in fnode.args.kwonlyargs]
        ours = list(sig.parameters.values())
        our_names = [p.name for p in ours]
        has_kwargs = any(p.kind == p.VAR_KEYWORD for p in ours)
        has_var = any(p.kind == p.VAR_POSITIONAL for p in ours)
        if names and names[0] in ("self", "cls"):
            names = names[1:]
        miss = [n for n in names + kwonly if n not in our_names and not has_kwargs]
        pos = [p.name for p in ours if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) and p.name not in ("self", "cls")]
        order_bad = (not has_var) and any(i >= len(pos) or pos[i] != n for i, n in enumerate(names) if n in our_names)
        if miss or order_bad:
            issues.append((label, miss, names, pos))

    for root, _, files in os.walk("/root/reference"):
        rel = os.path.relpath(root, "/root/reference")
        if rel.split(os.sep)[0] in ("tests", "examples", ".git"):
            continue
        for f in files:
            if not f.endswith(".py") or f in ("setup.py", "test.py", "test_weight_shapes.py") or "modal_run" in f:
                continue
            mod = os.path.join(rel, f[:-3]).replace(os.sep, ".").lstrip(".")
            mod = mod[:-9] if mod.endswith(".__init__") else mod
            mod = "" if mod in ("__init__", "") else mod
            target = importlib.import_module("QuintNet" + ("." + mod if mod else ""))
            for n in ast.parse(open(os.path.join(root, f)).read()).body:
                if isinstance(n, ast.FunctionDef) and not n.name.startswith("_"):
                    check(f"{mod}.{n.name}", n, getattr(target, n.name))
                if isinstance(n, ast.ClassDef) and not n.name.startswith("_"):
                    cls = getattr(target, n.name)
                    for b in n.body:
                        if isinstance(b, ast.FunctionDef) and (not b.name.startswith("_") or b.name == "__init__"):
                            if b.name in ("forward", "backward") and any(
                                    isinstance(d, ast.Name) and d.id == "staticmethod" for d in b.decorator_list):
                                continue                      # autograd.Function statics: ctx-based, not a user API
                            obj = getattr(cls, b.name, None)
                            if obj is not None:
                                check(f"{mod}.{n.name}.{b.name}", b, obj)
    assert not issues, issues


def test_reference_style_helper_calls(tmp_path):
    """The helpers whose signatures were aligned with the reference behave: utils.utils / utils.checkpoint save+load,
    Logger(filename, stream), setup_logger(log_file=…), merge helpers with payload dicts, distribute_layers(total_layers=…)."""
    import io
    import torch.nn as nn
    from QuintNet.parallelism.pipeline_parallel.utils import distribute_layers
    from QuintNet.utils import checkpoint as ck
    from QuintNet.utils import utils as uu
    from QuintNet.utils.logger import Logger
    from QuintNet.utils.logging import log_rank_0, setup_logger
    from QuintNet.merge_checkpoints import convert_to_hf_format, merge_pp_stages, merge_tp_shards
    m, opt = nn.Linear(4, 3), None
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    uu.save_checkpoint(m, opt, 3, 0.5, str(tmp_path / "a.pt"))
    m2 = nn.Linear(4, 3)
    _, _, epoch, loss = uu.load_checkpoint(m2, torch.optim.SGD(m2.parameters(), lr=0.1), str(tmp_path / "a.pt"))
    assert epoch == 3 and loss == 0.5 and torch.equal(m.weight, m2.weight) and uu.count_parameters(m) == 15
    path = ck.save_checkpoint(m, opt, epoch=4, checkpoint_dir=tmp_path / "d", is_best=True, loss=0.25)
    assert os.path.exists(path) and os.path.exists(tmp_path / "d" / "best_model.pt")
    m3 = nn.Linear(4, 3)
    meta = ck.load_checkpoint(checkpoint_path=path, model=m3, optimizer=None, strict=True)
    assert meta["epoch"] == 4 and meta["loss"] == 0.25 and torch.equal(m.weight, m3.weight)
    buf = io.StringIO()
    lg = Logger(str(tmp_path / "tee.log"), buf)
    lg.write("hello"); lg.flush()
    assert buf.getvalue() == "hello" and open(tmp_path / "tee.log").read() == "hello"
    logger = setup_logger(name="qn_test", level="INFO", log_file=str(tmp_path / "log" / "x.log"))
    log_rank_0("visible", logger=logger)
    for h in logger.handlers:
        h.flush()
    assert "visible" in open(tmp_path / "log" / "x.log").read()
    assert distribute_layers(total_layers=5, num_stages=2) == [[0, 1, 2], [3, 4]]
    tp = {t: {"model_state_dict": {"blocks.0.mlp.c_fc.proj.weight": torch.full((2, 3), float(t)),
                                     "blocks.0.ln_1.weight": torch.ones(3)}} for t in range(2)}
    stage = merge_tp_shards(tp_shards=tp, pp_rank=0)
    assert stage["blocks.0.mlp.c_fc.proj.weight"].shape == (4, 3)
    merged = merge_pp_stages(pp_stages={0: stage, 1: dict(stage)}, config={"n_layer": 2})
    assert {"blocks.0.ln_1.weight", "blocks.1.ln_1.weight"} <= set(merged)
    assert "transformer.h.1.mlp.c_fc.weight" in convert_to_hf_format(merged_state=merged)
