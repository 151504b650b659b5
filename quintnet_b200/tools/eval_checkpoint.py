"""Single-device evaluation of a pretrained / merged GPT-2 checkpoint: token-level loss & perplexity (with the
next-token shift) and optional ROUGE/BLEU on generated summaries — the role of the reference's ``test.py``
(root, 387 LoC), without the HuggingFace model dependency: the checkpoint is loaded into
:class:`quintnet_b200.models.gpt2.GPT2Model` through the sharded loader (tp = pp = 1).

    python -m quintnet_b200.tools.eval_checkpoint --checkpoint merged.pt [--dataset dir --tokenizer gpt2]
"""
from __future__ import annotations

import argparse
import math

import torch

from ..core.distributed_loading import shard_hf_tensors
from ..models.gpt2 import GPT2Config, GPT2Stage
from ..utils.data import SyntheticTokens, dict_collate


def load_model(path: str, device, dtype=None):
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        sd, cfg = load_file(path), None
    else:
        ck = torch.load(path, map_location="cpu", weights_only=False)
        sd, cfg = ck.get("model_state_dict", ck), ck.get("config")
    full = {k.replace("transformer.", ""): v for k, v in sd.items()}
    if not cfg:
        E = full["wte.weight"].shape[1]
        cfg = dict(vocab_size=full["wte.weight"].shape[0], n_positions=full["wpe.weight"].shape[0], n_embd=E,
                   n_layer=1 + max(int(k.split(".")[1]) for k in full if k.startswith("h.")), n_head=max(E // 64, 1))
    config = GPT2Config.from_dict(dict(cfg, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0))
    local = shard_hf_tensors(lambda n: full[n], config, 0, 1, 0, 1, device, dtype)
    model = GPT2Stage.from_sharded_state_dict(local, config, 0, 1, 0, 1, None, None, device, dtype=dtype).eval()
    return model, config


@torch.no_grad()
def evaluate_loss(model, loader, device):
    tot, n = 0.0, 0.0
    for batch in loader:
        ls, k = model(batch["input_ids"].to(device), labels=batch["labels"].to(device))
        tot, n = tot + float(ls), n + float(k)
    loss = tot / max(n, 1.0)
    return loss, (math.exp(loss) if loss < 20 else float("inf"))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--dataset", default=None, help="summarisation CSV dir (validation.csv); synthetic tokens if omitted")
    ap.add_argument("--tokenizer", default="gpt2")
    ap.add_argument("--max_samples", type=int, default=256)
    ap.add_argument("--generation", action="store_true", help="also compute ROUGE/BLEU on greedy summaries")
    a = ap.parse_args(argv)
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    model, config = load_model(a.checkpoint, device, torch.bfloat16 if device.type == "cuda" else None)
    tok = None
    if a.dataset:
        from transformers import GPT2Tokenizer
        from ..utils.data import SummarizationCollator, SummarizationDataset
        tok = GPT2Tokenizer.from_pretrained(a.tokenizer)
        ds = SummarizationDataset(a.dataset, "validation", a.max_samples)
        loader = torch.utils.data.DataLoader(ds, batch_size=8, collate_fn=SummarizationCollator(tok, 512))
    else:
        ds = SyntheticTokens(a.max_samples, min(512, config.n_positions), config.vocab_size, seed=1, min_fill=0.5)
        loader = torch.utils.data.DataLoader(ds, batch_size=8, collate_fn=dict_collate)
    loss, ppl = evaluate_loss(model, loader, device)
    print(f"loss {loss:.4f} | perplexity {ppl:.2f} over {len(ds)} samples")
    if a.generation and tok is not None:
        from ..utils.metrics import evaluate_generation
        print(evaluate_generation(model, tok, ds, device, num_samples=min(50, len(ds))))


if __name__ == "__main__":
    main()
