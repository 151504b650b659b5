"""GPT-2 summarisation fine-tune under 3-D parallelism (reference ``examples/gpt2_finetune.py:37-250``).

    torchrun --nproc-per-node 8 -m examples.gpt2_finetune --config examples/gpt2_config.yaml \
        --checkpoint /path/model.safetensors --dataset /path/cnn_dailymail --tokenizer gpt2
Without ``--dataset`` (or with ``--synthetic``) random token sequences of the configured length are used and
without ``--checkpoint`` the weights are a seeded random init — the path the benchmarks take.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quintnet_b200 import GPT2Trainer, cleanup_distributed, get_strategy, init_process_groups, load_config, setup_distributed
from quintnet_b200.models.gpt2 import GPT2Config
from quintnet_b200.utils.data import (SummarizationCollator, SummarizationDataset, SyntheticTokens, make_loader)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=os.path.join(os.path.dirname(__file__), "gpt2_config.yaml"))
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--dataset", default=None)
    ap.add_argument("--tokenizer", default="gpt2")
    ap.add_argument("--max_samples", type=int, default=None)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    config = load_config(a.config)
    if a.cpu:
        config["device_type"], config["dtype"] = "cpu", "float32"
    setup_distributed(device_type=config.get("device_type", "cuda"))
    pg = init_process_groups(config.get("device_type", "cuda"), config["mesh_dim"], config["mesh_name"])
    pg.print_mesh_info()
    model_config = GPT2Config.from_dict(config["model_config"])
    dp = pg.axis_size("dp")
    mb = max(int(config["batch_size"]) // (int(config["grad_acc_steps"]) * dp), 1)
    T = int(config.get("max_seq_length", 512))
    n_train = a.max_samples or int(config.get("max_samples", 2048))
    n_val = int(config.get("max_val_samples", 256))
    tokenizer = None
    if a.dataset and not a.synthetic:
        from transformers import GPT2Tokenizer
        tokenizer = GPT2Tokenizer.from_pretrained(a.tokenizer)
        tokenizer.pad_token = tokenizer.eos_token
        coll = SummarizationCollator(tokenizer, T, int(config.get("max_target_length", 128)))
        from torch.utils.data import DataLoader
        from torch.utils.data.distributed import DistributedSampler
        tr, va = SummarizationDataset(a.dataset, "train", n_train), SummarizationDataset(a.dataset, "validation", n_val)
        mk = lambda ds, sh: DataLoader(ds, batch_size=mb, collate_fn=coll, drop_last=True,
                                       sampler=DistributedSampler(ds, dp, pg.axis_rank("dp"), shuffle=sh))
        train_loader, val_loader = mk(tr, True), mk(va, False)
    else:
        train_loader = make_loader(SyntheticTokens(n_train, T, model_config.vocab_size, seed=0, min_fill=0.5), mb, pg)
        val_loader = make_loader(SyntheticTokens(n_val, T, model_config.vocab_size, seed=1, min_fill=0.5), mb, pg, shuffle=False)
    model = get_strategy("3d", pg, config, checkpoint_path=a.checkpoint, is_staged=True).apply(None)
    GPT2Trainer(model, train_loader, val_loader, config, pg, model_config, tokenizer=tokenizer).fit()
    cleanup_distributed()


if __name__ == "__main__":
    main()
