"""ROUGE / BLEU + greedy generation (reference ``utils/metrics.py:12-206``).
``rouge_score`` / ``sacrebleu`` are optional; without them a built-in
n-gram ROUGE-N / ROUGE-L / BLEU implementation is used (the reference
soft-fails to zeros)."""
from __future__ import annotations

import math
from collections import Counter
from typing import Dict, List

import torch


def _ngrams(tokens: List[str], n: int) -> Counter:
    return Counter(tuple(tokens[i:i + n]) for i in range(len(tokens) - n + 1))


def _rouge_n(pred: List[str], ref: List[str], n: int) -> float:
    p, r = _ngrams(pred, n), _ngrams(ref, n)
    if not p or not r:
        return 0.0
    overlap = sum((p & r).values())
    prec, rec = overlap / sum(p.values()), overlap / sum(r.values())
    return 0.0 if prec + rec == 0 else 2 * prec * rec / (prec + rec)


def _lcs(a: List[str], b: List[str]) -> int:
    dp = [0] * (len(b) + 1)
    for x in a:
        prev = 0
        for j, y in enumerate(b, 1):
            cur = dp[j]
            dp[j] = prev + 1 if x == y else max(dp[j], dp[j - 1])
            prev = cur
    return dp[-1]


def _rouge_l(pred: List[str], ref: List[str]) -> float:
    if not pred or not ref:
        return 0.0
    l = _lcs(pred, ref)
    prec, rec = l / len(pred), l / len(ref)
    return 0.0 if prec + rec == 0 else 2 * prec * rec / (prec + rec)


def _bleu(preds: List[List[str]], refs: List[List[str]], max_n: int = 4) -> float:
    log_p, plen, rlen = 0.0, 0, 0
    for n in range(1, max_n + 1):
        match, total = 0, 0
        for p, r in zip(preds, refs):
            pc, rc = _ngrams(p, n), _ngrams(r, n)
            match += sum((pc & rc).values())
            total += max(sum(pc.values()), 0)
        log_p += math.log((match + 1e-9) / (total + 1e-9)) / max_n
    for p, r in zip(preds, refs):
        plen, rlen = plen + len(p), rlen + len(r)
    bp = 1.0 if plen > rlen else math.exp(1 - rlen / max(plen, 1))
    return 100.0 * bp * math.exp(log_p)


def compute_rouge_bleu(predictions: List[str], references: List[str]) -> Dict[str, float]:
    try:
        from rouge_score import rouge_scorer
        import sacrebleu
        sc = rouge_scorer.RougeScorer(["rouge1", "rouge2", "rougeL"], use_stemmer=True)
        r1 = r2 = rl = 0.0
        for p, r in zip(predictions, references):
            s = sc.score(r, p)
            r1, r2, rl = r1 + s["rouge1"].fmeasure, r2 + s["rouge2"].fmeasure, rl + s["rougeL"].fmeasure
        n = max(len(predictions), 1)
        bleu = sacrebleu.corpus_bleu(predictions, [references]).score
        return {"rouge1": r1 / n, "rouge2": r2 / n, "rougeL": rl / n, "bleu": bleu}
    except ImportError:
        pt = [p.lower().split() for p in predictions]
        rt = [r.lower().split() for r in references]
        n = max(len(pt), 1)
        return {"rouge1": sum(_rouge_n(p, r, 1) for p, r in zip(pt, rt)) / n,
                "rouge2": sum(_rouge_n(p, r, 2) for p, r in zip(pt, rt)) / n,
                "rougeL": sum(_rouge_l(p, r) for p, r in zip(pt, rt)) / n,
                "bleu": _bleu(pt, rt)}


@torch.no_grad()
def generate_summary(model, tokenizer, input_text: str, max_new_tokens: int = 128, device=None,
                     max_input_length: int = 384) -> str:
    """Greedy summary of ``input_text`` (reference signature ``utils/metrics.py:74-80``)."""
    if device is None:
        device = next(model.parameters()).device
    prompt = f"{input_text}\n\nTL;DR: "
    ids = tokenizer(prompt, return_tensors="pt", truncation=True, max_length=max_input_length)["input_ids"].to(device)
    model.eval()
    out = ids
    n_pos = getattr(getattr(model, "config", None), "n_positions", 1024)
    for _ in range(max_new_tokens):
        logits = model(out[:, -n_pos:])
        nxt = logits[:, -1].float().argmax(-1, keepdim=True)
        out = torch.cat([out, nxt], dim=1)
        if tokenizer.eos_token_id is not None and int(nxt) == tokenizer.eos_token_id:
            break
    return tokenizer.decode(out[0, ids.shape[1]:], skip_special_tokens=True)


def evaluate_generation(model, tokenizer, val_dataset, device, num_samples: int = 50,
                        max_new_tokens: int = 128) -> Dict[str, float]:
    dataset = val_dataset
    preds, refs = [], []
    for i in range(min(num_samples, len(dataset))):
        item = dataset[i]
        preds.append(generate_summary(model, tokenizer, item["article"], max_new_tokens, device))
        refs.append(item["highlights"])
    return compute_rouge_bleu(preds, refs)
