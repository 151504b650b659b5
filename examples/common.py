"""Shared plumbing for the example entry points (reference ``examples/*.py``)."""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from quintnet_b200 import Trainer, get_strategy, init_process_groups, load_config, setup_distributed, cleanup_distributed
from quintnet_b200.models import Model
from quintnet_b200.utils.data import CustomDataset, SyntheticMNIST, make_loader, mnist_transform


def base_parser(default_config: str) -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=os.path.join(os.path.dirname(__file__), default_config))
    ap.add_argument("--synthetic", action="store_true", help="MNIST-shaped synthetic data (no dataset needed)")
    ap.add_argument("--cpu", action="store_true", help="run on CPU / gloo")
    ap.add_argument("--epochs", type=int, default=None)
    ap.add_argument("--samples", type=int, default=2048, help="synthetic dataset size")
    return ap


def vit_loaders(config, pg, synthetic: bool, n: int):
    dp = pg.axis_size("dp")
    mb = max(int(config["batch_size"]) // (int(config.get("grad_acc_steps", 1)) * dp), 1)
    path = config.get("dataset_path", "")
    if synthetic or not os.path.exists(path):
        train, val = SyntheticMNIST(n, seed=0), SyntheticMNIST(max(n // 8, 64), seed=1)
    else:
        train = CustomDataset(path, "train", mnist_transform)
        val = CustomDataset(path, "test", mnist_transform)
    return (make_loader(train, mb, pg, shuffle=True, num_workers=int(config.get("num_workers", 0))),
            make_loader(val, mb, pg, shuffle=False))


def run_vit(strategy_name: str, default_config: str):
    args = base_parser(default_config).parse_args()
    config = load_config(args.config)
    config["strategy_name"] = strategy_name or config.get("strategy_name", "3d")
    if args.cpu:
        config["device_type"] = "cpu"
    if args.epochs:
        config["num_epochs"] = args.epochs
    setup_distributed(device_type=config.get("device_type", "cuda"))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if len(config["mesh_dim"]) == 1 and int(config["mesh_dim"][0]) != world:
        print(f"[examples] 1-D mesh {config['mesh_dim']} resized to the launch size [{world}]")
        config["mesh_dim"] = [world]
    pg = init_process_groups(config.get("device_type", "cuda"), config["mesh_dim"], config["mesh_name"])
    pg.print_mesh_info()
    train_loader, val_loader = vit_loaders(config, pg, args.synthetic, args.samples)
    model = Model(img_size=config["img_size"], patch_size=config["patch_size"], hidden_dim=config["hidden_dim"],
                  in_channels=config["in_channels"], n_heads=config["n_heads"], depth=config["depth"])
    model = get_strategy(config["strategy_name"], pg, config).apply(model)
    trainer = Trainer(model, train_loader, val_loader, config, pg)
    t0 = time.time()
    trainer.fit()
    if pg.rank == 0:
        print(f"Total training time: {time.time() - t0:.2f} seconds")
    cleanup_distributed()
