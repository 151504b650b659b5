"""Datasets, collators and synthetic stand-ins (reference ``utils/Dataloader.py:38-358``).

``CustomDataset`` / ``SummarizationDataset`` / ``SummarizationCollator`` /
``SummarizationDataLoader`` keep their reference signatures.  The synthetic
datasets produce tensors of exactly the shapes the real pipelines produce
(MNIST ``[1,28,28]`` + label; GPT-2 ``input_ids/attention_mask/labels`` padded
to ``max_length``) — the GPU box has no network, so benchmarks and tests run on
these (BASELINE.json: "synthetic data of the named shape").
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional

import torch
from torch.utils.data import DataLoader, Dataset


# ---------------------------------------------------------------- real datasets
class CustomDataset(Dataset):
    """HF ``datasets`` directory (``load_from_disk``) or ``.arrow`` file with
    ``image`` / ``label`` columns."""

    def __init__(self, dataset_path: str, split: str = "train", transform: Optional[Callable] = None):
        try:
            import datasets  # noqa: F401
        except ImportError as e:  # pragma: no cover
            raise ImportError("CustomDataset needs the HuggingFace `datasets` package") from e
        from datasets import Dataset as HFDataset, load_from_disk
        self.transform = transform
        self.dataset_path, self.split = dataset_path, split
        if os.path.isdir(dataset_path):
            ds = load_from_disk(dataset_path)
            self.data = ds[split] if hasattr(ds, "keys") and split in ds else ds
        elif dataset_path.endswith(".arrow"):
            self.data = HFDataset.from_file(dataset_path)
        else:
            raise ValueError(f"unsupported dataset path {dataset_path}")

    @property
    def dataset(self):                # reference attribute name
        return self.data

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        item = self.data[idx]
        sample = {"image": item["image"], "label": item["label"]}
        return self.transform(sample) if self.transform else sample


def pil_to_tensor_transform(sample):
    import numpy as np
    img = torch.from_numpy(np.array(sample["image"], dtype="float32") / 255.0)
    if img.dim() == 2:
        img = img.unsqueeze(0)
    elif img.dim() == 3:
        img = img.permute(2, 0, 1)
    return {"image": img, "label": torch.tensor(sample["label"], dtype=torch.long)}


def mnist_transform(sample):
    out = pil_to_tensor_transform(sample)
    out["image"] = (out["image"] - 0.1307) / 0.3081
    return out


class SummarizationDataset(Dataset):
    """CSV with ``article,highlights`` columns: ``{train,validation,test}.csv``."""

    def __init__(self, dataset_path: str, split: str = "train", max_samples: Optional[int] = None):
        import pandas as pd
        path = os.path.join(dataset_path, f"{split}.csv") if os.path.isdir(dataset_path) else dataset_path
        df = pd.read_csv(path)
        if max_samples:
            df = df.iloc[:max_samples]
        self.dataset_path, self.split, self.data = dataset_path, split, df
        self.articles: List[str] = df["article"].astype(str).tolist()
        self.highlights: List[str] = df["highlights"].astype(str).tolist()

    def __len__(self):
        return len(self.articles)

    def __getitem__(self, idx):
        return {"article": self.articles[idx], "highlights": self.highlights[idx]}


class SummarizationCollator:
    """``"{article}\\n\\nTL;DR: {highlights}{eos}"`` padded to ``max_length``;
    labels = input_ids with pad → −100 (``mask_prompt=True`` additionally masks
    the article, which the reference's docstring promises but does not do)."""

    def __init__(self, tokenizer, max_length: int = 512, max_target_length: int = 128,
                 mask_prompt: bool = False):
        self.tokenizer, self.max_length = tokenizer, max_length
        self.max_target_length, self.mask_prompt = max_target_length, mask_prompt
        if getattr(tokenizer, "pad_token", None) is None:
            tokenizer.pad_token = tokenizer.eos_token

    def __call__(self, batch: List[Dict[str, str]]) -> Dict[str, torch.Tensor]:
        eos = self.tokenizer.eos_token or ""
        prompts = [f"{b['article']}\n\nTL;DR: " for b in batch]
        texts = [p + f"{b['highlights']}{eos}" for p, b in zip(prompts, batch)]
        enc = self.tokenizer(texts, max_length=self.max_length, truncation=True, padding="max_length",
                             return_tensors="pt")
        labels = enc["input_ids"].clone()
        labels[enc["attention_mask"] == 0] = -100
        if self.mask_prompt:
            for i, p in enumerate(prompts):
                n = len(self.tokenizer(p, truncation=True, max_length=self.max_length)["input_ids"])
                labels[i, :n] = -100
        return {"input_ids": enc["input_ids"], "attention_mask": enc["attention_mask"], "labels": labels}


class SummarizationDataLoader(DataLoader):
    """Reference signature ``(dataset, batch_size, collator, shuffle=True, num_workers=0)``
    (``utils/Dataloader.py:334-341``); without a ``collator`` one is built from ``tokenizer`` / ``max_length`` /
    ``max_target_length``."""

    def __init__(self, dataset, batch_size: int = 8, collator=None, shuffle: bool = True, num_workers: int = 0,
                 tokenizer=None, max_length: int = 512, max_target_length: int = 128, sampler=None, **kw):
        if collator is None:
            if tokenizer is None:
                raise ValueError("SummarizationDataLoader needs a collator or a tokenizer")
            collator = SummarizationCollator(tokenizer, max_length, max_target_length)
        self.collator = collator
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle if sampler is None else False,
                         sampler=sampler, num_workers=num_workers, collate_fn=collator, **kw)

    @property
    def data_loader(self):          # the reference wraps a DataLoader in this attribute
        return self


# ------------------------------------------------------------------- synthetic
class SyntheticMNIST(Dataset):
    """Learnable MNIST-shaped data: class-dependent blob pattern + noise."""

    def __init__(self, n: int = 1024, num_classes: int = 10, img_size: int = 28, channels: int = 1,
                 seed: int = 0, noise: float = 0.5):
        g = torch.Generator().manual_seed(seed)
        self.labels = torch.randint(0, num_classes, (n,), generator=g)
        proto = torch.randn(num_classes, channels, img_size, img_size,
                            generator=torch.Generator().manual_seed(1234))
        self.images = proto[self.labels] + noise * torch.randn(n, channels, img_size, img_size, generator=g)

    def __len__(self):
        return self.labels.shape[0]

    def __getitem__(self, i):
        return {"image": self.images[i], "label": self.labels[i]}


class SyntheticTokens(Dataset):
    """Random token sequences of fixed length, right-padded like the
    summarisation collator's output."""

    def __init__(self, n: int, seq_len: int = 512, vocab_size: int = 50257, seed: int = 0,
                 min_fill: float = 1.0, pad_id: int = 50256):
        g = torch.Generator().manual_seed(seed)
        self.input_ids = torch.randint(0, vocab_size, (n, seq_len), generator=g)
        lens = (torch.rand(n, generator=g) * (1.0 - min_fill) + min_fill) * seq_len
        lens = lens.long().clamp(2, seq_len)
        pos = torch.arange(seq_len)[None, :]
        self.attention_mask = (pos < lens[:, None]).long()
        self.input_ids[self.attention_mask == 0] = min(pad_id, vocab_size - 1)
        self.labels = self.input_ids.clone()
        self.labels[self.attention_mask == 0] = -100

    def __len__(self):
        return self.input_ids.shape[0]

    def __getitem__(self, i):
        return {"input_ids": self.input_ids[i], "attention_mask": self.attention_mask[i],
                "labels": self.labels[i]}


def dict_collate(batch):
    return {k: torch.stack([b[k] for b in batch]) for k in batch[0]}


def make_loader(dataset, batch_size: int, pg_manager=None, shuffle: bool = True, drop_last: bool = True,
                num_workers: int = 0, pin_memory: bool = False, seed: int = 0) -> DataLoader:
    """DataLoader whose sampler shards by the *dp* axis (same data on all tp/pp
    ranks of a replica — what the pipeline schedules require)."""
    from torch.utils.data.distributed import DistributedSampler
    dp_size = pg_manager.axis_size("dp") if pg_manager is not None else 1
    dp_rank = pg_manager.axis_rank("dp") if pg_manager is not None else 0
    sampler = DistributedSampler(dataset, num_replicas=dp_size, rank=dp_rank, shuffle=shuffle,
                                 seed=seed, drop_last=drop_last)
    return DataLoader(dataset, batch_size=batch_size, sampler=sampler, drop_last=drop_last,
                      num_workers=num_workers, pin_memory=pin_memory, collate_fn=dict_collate)
