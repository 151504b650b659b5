"""Reload ``model_checkpoint.pt`` into a fresh single-device ViT and re-measure accuracy
(reference ``examples/verify_model.py:23-131``)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quintnet_b200.models import Model
from quintnet_b200.utils.data import CustomDataset, SyntheticMNIST, dict_collate, mnist_transform


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint", default="model_checkpoint.pt")
    ap.add_argument("--dataset", default=None)
    ap.add_argument("--synthetic", action="store_true")
    a = ap.parse_args()
    ck = torch.load(a.checkpoint, map_location="cpu", weights_only=False)
    cfg = ck["config"]
    model = Model(cfg["img_size"], cfg["patch_size"], cfg["hidden_dim"], cfg["in_channels"], cfg["n_heads"], cfg["depth"])
    sd = ck["model_state_dict"]
    if "pipeline_stage" in sd:
        raise SystemExit("checkpoint holds a single pipeline stage; verify a non-PP (dp/tp) run")
    missing, unexpected = model.load_state_dict(sd, strict=False)
    print(f"loaded ({len(missing)} missing, {len(unexpected)} unexpected keys)")
    data = SyntheticMNIST(1024, seed=1) if (a.synthetic or not a.dataset) else CustomDataset(a.dataset, "test", mnist_transform)
    loader = torch.utils.data.DataLoader(data, batch_size=256, collate_fn=dict_collate)
    model.eval()
    correct = total = 0
    with torch.no_grad():
        for b in loader:
            correct += (model(b["image"]).argmax(1) == b["label"]).sum().item()
            total += b["label"].numel()
    print(f"accuracy: {100.0 * correct / total:.2f}% ({correct}/{total})")


if __name__ == "__main__":
    main()
