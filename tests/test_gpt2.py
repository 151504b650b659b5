"""GPT-2: schema, TP (plain + sequence-parallel) == single device, PP schedules ==
single process (loss AND post-step parameters), 3-D 2x2x2 parity."""
import copy
import os

import pytest

import torch
import torch.distributed as dist
import torch.nn as nn

from tests._dist import run_distributed
from quintnet_b200 import init_process_groups, get_strategy, GPT2Trainer
from quintnet_b200.models.gpt2 import GPT2Config, GPT2Model, GPT2Stage
from quintnet_b200.core.distributed_loading import random_gpt2_state_dict
from quintnet_b200.utils.data import SyntheticTokens, make_loader

CFG = dict(vocab_size=96, n_positions=16, n_embd=32, n_layer=4, n_head=4,
           resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)


def test_state_dict_schema():
    cfg = GPT2Config(**CFG)
    m = GPT2Model(cfg)
    keys = set(m.state_dict().keys())
    for k in ["embedding.wte.weight", "embedding.wpe.weight", "blocks.0.ln_1.weight",
              "blocks.0.attn.c_attn.proj.weight", "blocks.0.attn.c_attn.proj.bias",
              "blocks.0.attn.c_proj.proj.weight", "blocks.0.attn.c_proj.bias",
              "blocks.3.mlp.c_fc.proj.weight", "blocks.3.mlp.c_proj.proj.weight",
              "blocks.3.mlp.c_proj.bias", "ln_f.weight", "ln_f.bias"]:
        assert k in keys, k
    assert m.state_dict()["blocks.0.attn.c_attn.proj.weight"].shape == (96, 32)
    ids = torch.randint(0, 96, (2, 16))
    logits = m(ids)
    assert logits.shape == (2, 16, 96)
    # fused loss == CE on shifted logits
    ls, n = m(ids, labels=ids)
    ref = nn.functional.cross_entropy(logits[:, :-1].reshape(-1, 96), ids[:, 1:].reshape(-1), reduction="sum")
    assert torch.allclose(ls, ref, atol=1e-4) and int(n) == 2 * 15
    out = m.generate(ids[:, :4], max_new_tokens=3)
    assert out.shape == (2, 7)


def test_hf_equivalence():
    """Our GPT-2 computes the HuggingFace GPT-2 function (exact-erf GELU)."""
    transformers = __import__("pytest").importorskip("transformers")
    from quintnet_b200.core.distributed_loading import hf_shapes, random_hf_tensor
    cfg = GPT2Config(**CFG)
    hf_cfg = transformers.GPT2Config(vocab_size=96, n_positions=16, n_embd=32, n_layer=4, n_head=4,
                                     resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0,
                                     activation_function="gelu")
    hf = transformers.GPT2LMHeadModel(hf_cfg).eval()
    sd = {("transformer." + n): random_hf_tensor(n, s, cfg, 0, torch.device("cpu"))
          for n, s in hf_shapes(cfg).items()}
    sd["lm_head.weight"] = sd["transformer.wte.weight"]
    hf.load_state_dict(sd, strict=False)
    ours = GPT2Model(cfg, seed=0).eval()
    ids = torch.randint(0, 96, (2, 16))
    with torch.no_grad():
        assert torch.allclose(ours(ids), hf(ids).logits, atol=2e-4)


def _tp_equals_single(rank, world):
    pg = init_process_groups("cpu", [world], ["tp"])
    g = pg.get_group("tp")
    for sp in (False, True):
        cfg = GPT2Config(**CFG, sequence_parallel=sp)
        ref = GPT2Model(GPT2Config(**CFG), seed=1)
        tp = GPT2Model(cfg, tp_group=g, tp_rank=rank, tp_size=world, seed=1)
        ids = torch.randint(0, 96, (4, 16), generator=torch.Generator().manual_seed(3))
        assert torch.allclose(tp(ids), ref(ids), atol=1e-4), f"logits sp={sp}"
        ls_r, n_r = ref(ids, labels=ids)
        ls_t, n_t = tp(ids, labels=ids)
        assert torch.allclose(ls_t, ls_r, atol=1e-3) and float(n_t) == float(n_r)
        (ls_r / n_r).backward()
        (ls_t / n_t).backward()
        tp.sync_sequence_parallel_grads()
        assert torch.allclose(tp.embedding.wte.weight.grad, ref.embedding.wte.weight.grad, atol=1e-4), sp
        assert torch.allclose(tp.blocks[0].ln_1.weight.grad, ref.blocks[0].ln_1.weight.grad, atol=1e-4), sp
        E = 32
        w = E // world
        cols = torch.cat([torch.arange(rank * w, (rank + 1) * w) + j * E for j in range(3)])
        assert torch.allclose(tp.blocks[1].attn.c_attn.proj.weight.grad,
                              ref.blocks[1].attn.c_attn.proj.weight.grad[cols], atol=1e-4), sp


@pytest.mark.parametrize("world", [2, 4])
def test_tp_equals_single(world):
    run_distributed(_tp_equals_single, world)


def _train_cfg(mesh_dim, mesh_name, schedule="1f1b", zero=0, sp=False, M=4, batch=8):
    return {"batch_size": batch, "learning_rate": 1e-2, "num_epochs": 1, "grad_acc_steps": M,
            "max_grad_norm": 1.0, "model_config": dict(CFG), "device_type": "cpu",
            "mesh_dim": mesh_dim, "mesh_name": mesh_name, "strategy_name": "3d", "schedule": schedule,
            "max_seq_length": 16, "zero_stage": zero, "sequence_parallel": sp, "seed": 1,
            "weight_decay": 0.01, "output_dir": "/tmp/qn_test_ckpt"}


def _single_process_reference(steps, M, batch, data):
    """Plain PyTorch loop: same model, same data order, AdamW, global clip."""
    cfg = GPT2Config(**CFG)
    model = GPT2Model(cfg, seed=1)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=0.01)
    losses = []
    idx = 0
    mb = batch // M
    for _ in range(steps):
        tot, cnt = 0.0, 0
        for _ in range(M):
            ids = data.input_ids[idx:idx + mb]
            lab = data.labels[idx:idx + mb]
            idx += mb
            ls, n = model(ids, labels=lab)
            (ls / n / M).backward()
            tot, cnt = tot + float(ls), cnt + float(n)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
        losses.append(tot / cnt)
    return losses, model


def _run_parallel(rank, world, mesh_dim, mesh_name, schedule, zero, sp, batch=8):
    M, steps = 4, 3
    dp = mesh_dim[mesh_name.index("dp")] if "dp" in mesh_name else 1
    # with dp>1 per-replica token-mean losses only equal the global one without padding
    data = SyntheticTokens(batch * steps, seq_len=16, vocab_size=96, seed=11, pad_id=0,
                           min_fill=0.6 if dp == 1 else 1.0)
    ref_losses, ref_model = _single_process_reference(steps, M, batch, data)
    cfg = _train_cfg(mesh_dim, mesh_name, schedule, zero, sp, M, batch)
    pg = init_process_groups("cpu", mesh_dim, mesh_name)
    model = get_strategy("3d", pg, cfg, checkpoint_path=None, is_staged=True).apply(None)
    # data order: a global step consumes `batch` samples; dp replica d takes micro-batches
    # d, d+dp, ... of the step (so the union over replicas is the reference's batch)
    mb = batch // (M * dp)
    d = pg.axis_rank("dp")
    order = []
    for s in range(steps):
        for m in range(M):
            start = s * batch + (m * dp + d) * mb
            order.extend(range(start, start + mb))
    sub = torch.utils.data.Subset(data, order)
    from quintnet_b200.utils.data import dict_collate
    loader = torch.utils.data.DataLoader(sub, batch_size=mb, shuffle=False, collate_fn=dict_collate)
    tr = GPT2Trainer(model, loader, None, cfg, pg, GPT2Config(**CFG))
    got = []
    for s in range(steps):
        loss, ppl = tr.train_step()
        if loss is not None:
            got.append(loss)
    # the dp replicas see different halves; average their reported losses
    if got:
        t = torch.tensor(got)
        if dp > 1:
            dist.all_reduce(t, group=pg.get_group("dp"))
            t /= dp
        assert torch.allclose(t, torch.tensor(ref_losses), atol=2e-3), (t, ref_losses)
    # parameters after 3 optimizer steps
    stage = tr.gpt2_stage
    if stage.is_last_stage:
        assert torch.allclose(stage.ln_f.weight.float(), ref_model.ln_f.weight, atol=3e-3)
    if stage.is_first_stage:
        assert torch.allclose(stage.embedding.wpe.weight.float(), ref_model.embedding.wpe.weight, atol=3e-3)
    from quintnet_b200.parallel.pipeline_parallel.utils import distribute_layers
    pp = pg.axis_size("pp")
    lid = distribute_layers(4, pp)[pg.axis_rank("pp")][0]
    assert torch.allclose(stage.blocks[0].ln_2.bias.float(), ref_model.blocks[lid].ln_2.bias, atol=3e-3)


def test_pp2_1f1b():
    run_distributed(_run_parallel, 2, [2], ["pp"], "1f1b", 0, False)


def test_pp2_afab():
    run_distributed(_run_parallel, 2, [2], ["pp"], "afab", 0, False)


def test_pp4_1f1b():
    run_distributed(_run_parallel, 4, [4], ["pp"], "1f1b", 0, False)


def test_3d_2x2x2():
    run_distributed(_run_parallel, 8, [2, 2, 2], ["dp", "tp", "pp"], "1f1b", 0, False, timeout=400)


def test_3d_2x2x2_zero1_sequence_parallel():
    run_distributed(_run_parallel, 8, [2, 2, 2], ["dp", "tp", "pp"], "1f1b", 1, True, 16, timeout=400)


def _replay(rank, world, out_dir):
    """Deterministic replay: the same seed and data give bitwise-identical parameters twice."""
    from quintnet_b200.utils.data import dict_collate
    M, batch, steps = 4, 8, 2
    pg = init_process_groups("cpu", [2], ["pp"])
    digests = []
    for _ in range(2):
        data = SyntheticTokens(batch * steps, seq_len=16, vocab_size=96, seed=11, pad_id=0, min_fill=0.6)
        cfg = _train_cfg([2], ["pp"], "1f1b", 0, False, M, batch)
        cfg["model_config"] = dict(CFG, embd_pdrop=0.1, attn_pdrop=0.1, resid_pdrop=0.1)   # RNG in the loop too
        torch.manual_seed(7)
        model = get_strategy("3d", pg, cfg, checkpoint_path=None, is_staged=True).apply(None)
        loader = torch.utils.data.DataLoader(data, batch_size=batch // M, shuffle=False, collate_fn=dict_collate)
        tr = GPT2Trainer(model, loader, None, cfg, pg, GPT2Config(**cfg["model_config"]))
        for _ in range(steps):
            tr.train_step()
        digests.append(torch.cat([p.detach().float().flatten() for p in tr.gpt2_stage.parameters()]))
    assert torch.equal(digests[0], digests[1]), (digests[0] - digests[1]).abs().max()


def test_deterministic_replay(tmp_path):
    run_distributed(_replay, 2, str(tmp_path))


def test_fp8_linear_option_cpu():
    """fp8_linear: projections run on (emulated) e4m3 operands; the model still matches the bf16/fp32 one closely
    and trains."""
    torch.manual_seed(0)
    cfg8 = GPT2Config(**dict(CFG, fp8_linear=True))
    m8, m = GPT2Model(cfg8, seed=1), GPT2Model(GPT2Config(**CFG), seed=1)
    ids = torch.randint(1, 96, (4, 16))
    l8, n8 = m8(ids, labels=ids)
    l, n = m(ids, labels=ids)
    assert abs(float(l8 / n8) - float(l / n)) < 0.05 * float(l / n)
    (l8 / n8).backward()
    assert all(torch.isfinite(p.grad).all() for p in m8.parameters() if p.grad is not None)
    assert m8.blocks[0].mlp.c_fc.fp8 and m8.blocks[0].attn.c_proj.fp8


def test_interleaved_table_is_consistent():
    """Every unit runs exactly once, after its dependency, and every send has its receive in the same tick."""
    from quintnet_b200.parallel.pipeline_parallel.interleaved import bubble_fraction, build_table
    for pp, V, M in ((2, 2, 4), (2, 2, 8), (4, 2, 8), (2, 3, 6), (4, 3, 12)):
        table = build_table(pp, V, M)
        S = pp * V
        ran = {}
        for r in range(pp):
            for t, tick in enumerate(table[r]):
                if tick.unit:
                    kind, v, m = tick.unit
                    key = (kind, v * pp + r, m)
                    assert key not in ran
                    ran[key] = t
        assert len(ran) == 2 * S * M
        for (kind, s, m), t in ran.items():
            if kind == "F" and s > 0:
                assert ran[("F", s - 1, m)] < t
            if kind == "B":
                assert ran[("F", s, m)] < t and (s == S - 1 or ran[("B", s + 1, m)] < t)
        for t in range(len(table[0])):
            sends = sorted((r, peer, cons) for r in range(pp) for peer, cons in table[r][t].sends)
            recvs = sorted((peer, r, cons) for r in range(pp) for peer, cons in table[r][t].recvs)
            assert sends == recvs
            for r in range(pp):           # NCCL matches by order within a pair: at most one tensor per direction per tick
                peers = [p for p, _ in table[r][t].sends]
                assert len(peers) == len(set(peers))
        assert bubble_fraction(pp, V, M) < bubble_fraction(pp, 1, M) + 1e-9


def _run_interleaved(rank, world, zero, sp, mesh_dim, mesh_name, out_dir, batch=8):
    from quintnet_b200.utils.data import dict_collate
    M, steps, V = 4, 3, 2
    dp = mesh_dim[mesh_name.index("dp")] if "dp" in mesh_name else 1
    data = SyntheticTokens(batch * steps, seq_len=16, vocab_size=96, seed=11, pad_id=0,
                           min_fill=0.6 if dp == 1 else 1.0)
    ref_losses, ref_model = _single_process_reference(steps, M, batch, data)
    cfg = _train_cfg(mesh_dim, mesh_name, "1f1b", zero, sp, M, batch)
    cfg["virtual_pipeline_stages"] = V
    cfg["output_dir"] = out_dir
    pg = init_process_groups("cpu", mesh_dim, mesh_name)
    model = get_strategy("3d", pg, cfg, checkpoint_path=None, is_staged=True).apply(None)
    mb = batch // (M * dp)
    d = pg.axis_rank("dp")
    order = []
    for s in range(steps):
        for m in range(M):
            start = s * batch + (m * dp + d) * mb
            order.extend(range(start, start + mb))
    loader = torch.utils.data.DataLoader(torch.utils.data.Subset(data, order), batch_size=mb, shuffle=False,
                                         collate_fn=dict_collate)
    tr = GPT2Trainer(model, loader, None, cfg, pg, GPT2Config(**CFG))
    assert tr.pipeline_trainer.schedule_type == "1f1b_interleaved"
    got = []
    for s in range(steps):
        loss, ppl = tr.train_step()
        if loss is not None:
            got.append(loss)
    if got:
        t = torch.tensor(got)
        if dp > 1:
            dist.all_reduce(t, group=pg.get_group("dp"))
            t /= dp
        assert torch.allclose(t, torch.tensor(ref_losses), atol=2e-3), (t, ref_losses)
    chunks = tr.gpt2_stage.chunks
    pp, r = pg.axis_size("pp"), pg.axis_rank("pp")
    assert len(chunks) == V
    if r == pp - 1:
        assert torch.allclose(chunks[-1].ln_f.weight.float(), ref_model.ln_f.weight, atol=3e-3)
    if r == 0:
        assert torch.allclose(chunks[0].embedding.wpe.weight.float(), ref_model.embedding.wpe.weight, atol=3e-3)
        assert torch.allclose(chunks[0].embedding.wte.weight.float(), ref_model.embedding.wte.weight, atol=3e-3)
    for v, chunk in enumerate(chunks):          # 4 layers over 4 virtual stages: chunk v of rank r is layer v*pp + r
        lid = v * pp + r
        assert torch.allclose(chunk.blocks[0].ln_2.bias.float(), ref_model.blocks[lid].ln_2.bias, atol=3e-3)
        if pg.axis_size("tp") == 1:
            assert torch.allclose(chunk.blocks[0].mlp.c_fc.proj.bias.float(),
                                  ref_model.blocks[lid].mlp.c_fc.proj.bias.float(), atol=3e-3)
    # checkpoint: one shard per virtual stage, mergeable like a 4-stage pipeline; resume restores bitwise
    tr._save_checkpoint("vp.pt")
    dist.barrier()
    before = torch.cat([p.detach().float().flatten() for p in tr.gpt2_stage.parameters()])
    with torch.no_grad():
        for p in tr.gpt2_stage.parameters():
            p.add_(1.0)
    tr.resume(os.path.join(out_dir, "vp.pt"))
    after = torch.cat([p.detach().float().flatten() for p in tr.gpt2_stage.parameters()])
    assert torch.equal(before, after)
    if dist.get_rank() == 0:
        names = sorted(f for f in os.listdir(out_dir) if f.startswith("vp_pp") and f.endswith(".pt") and "opt" not in f)
        assert len(names) == pp * V * pg.axis_size("tp"), names


def test_pp2_interleaved_virtual_stages(tmp_path):
    run_distributed(_run_interleaved, 2, 0, False, [2], ["pp"], str(tmp_path))
    # the 2 ranks x 2 chunks wrote 4 stage shards: they merge like a 4-stage pipeline into the HF layout
    from quintnet_b200.tools.merge_checkpoints import merge_checkpoints
    steps, M, batch = 3, 4, 8
    data = SyntheticTokens(batch * steps, seq_len=16, vocab_size=96, seed=11, pad_id=0, min_fill=0.6)
    _, ref_model = _single_process_reference(steps, M, batch, data)
    hf = merge_checkpoints(str(tmp_path), str(tmp_path / "merged.pt"), "vp")
    for i in range(4):
        w = hf[f"transformer.h.{i}.mlp.c_fc.weight"]                     # HF Conv1D layout [in, out]
        assert torch.allclose(w.float(), ref_model.blocks[i].mlp.c_fc.proj.weight.t().float(), atol=3e-3), i
    assert torch.allclose(hf["transformer.wte.weight"].float(), ref_model.embedding.wte.weight.float(), atol=3e-3)


def _eval_interleaved(rank, world):
    """Forward-only (validation) pass of the virtual-stage pipeline == the single-process loss."""
    from quintnet_b200.utils.data import dict_collate
    data = SyntheticTokens(16, seq_len=16, vocab_size=96, seed=5, pad_id=0, min_fill=0.6)
    ref = GPT2Model(GPT2Config(**CFG), seed=1)
    cfg = _train_cfg([2], ["pp"], "1f1b", 0, False, 4, 8)
    cfg["virtual_pipeline_stages"] = 2
    pg = init_process_groups("cpu", [2], ["pp"])
    model = get_strategy("3d", pg, cfg, checkpoint_path=None, is_staged=True).apply(None)
    loader = torch.utils.data.DataLoader(data, batch_size=4, shuffle=False, collate_fn=dict_collate)
    tr = GPT2Trainer(model, loader, loader, cfg, pg, GPT2Config(**CFG))
    loss, ppl = tr.pipeline_trainer.evaluate(loader, tr._shapes(4), tr.device, tr.dtype)
    if pg.axis_rank("pp") == 1:
        with torch.no_grad():
            ls, n = ref(data.input_ids, labels=data.labels)
        assert abs(loss - float(ls / n)) < 2e-3, (loss, float(ls / n))
    else:
        assert loss is None


def test_pp2_interleaved_evaluate():
    run_distributed(_eval_interleaved, 2)


def test_3d_2x2x2_interleaved_zero1_sp(tmp_path):
    run_distributed(_run_interleaved, 8, 1, True, [2, 2, 2], ["dp", "tp", "pp"], str(tmp_path), 16, timeout=600)


def test_activation_checkpointing_matches():
    """Recomputing the blocks in backward gives the same loss and gradients, dropout included (RNG replay)."""
    cfg_d = dict(CFG, embd_pdrop=0.1, attn_pdrop=0.0, resid_pdrop=0.1)
    ids = torch.randint(1, 96, (4, 16))
    grads = []
    for ckpt in (False, True):
        model = GPT2Model(GPT2Config(**dict(cfg_d, activation_checkpointing=ckpt)), seed=1).train()
        torch.manual_seed(123)
        ls, n = model(ids, labels=ids)
        (ls / n).backward()
        grads.append((float(ls / n), torch.cat([p.grad.flatten() for p in model.parameters()])))
    assert abs(grads[0][0] - grads[1][0]) < 1e-6
    assert torch.allclose(grads[0][1], grads[1][1], atol=1e-6)


def test_3d_2x2x2_fused_wgrad_accumulation(monkeypatch):
    """Weight gradients added into the gradient buffers by the backward ops themselves (no AccumulateGrad add): same
    losses and parameters as the single-process loop under dp2 x tp2 x pp2 + ZeRO-1 + sequence parallelism."""
    monkeypatch.setenv("QUINTNET_FUSE_WGRAD_ACCUM", "1")
    run_distributed(_run_parallel, 8, [2, 2, 2], ["dp", "tp", "pp"], "1f1b", 1, True, 16, timeout=600)
    # plain DDP (per-bucket all-reduce driven by the post-accumulate hooks): the mode in which calling the hook by hand
    # would double-count and reduce buckets before their last gradient has arrived
    run_distributed(_run_parallel, 8, [2, 2, 2], ["dp", "tp", "pp"], "1f1b", 0, False, timeout=600)
