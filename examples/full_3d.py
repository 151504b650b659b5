"""3-D parallel training.  ViT on MNIST (default) or, with ``--staged``, build a GPT-2 stage per rank
from a checkpoint / random init (reference ``examples/full_3d.py:33-187``).

    torchrun --nproc-per-node 8 -m examples.full_3d --config examples/config.yaml [--synthetic]
    torchrun --nproc-per-node 8 -m examples.full_3d --staged --config examples/gpt2_config.yaml [--checkpoint model.safetensors]
"""
import sys

from examples.common import *  # noqa: F401,F403


def main():
    if "--staged" not in sys.argv:
        return run_vit("3d", "config.yaml")
    ap = base_parser("gpt2_config.yaml")
    ap.add_argument("--staged", action="store_true")
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--model_config", default=None, help="JSON file with the GPT-2 model config (overrides the YAML's)")
    args = ap.parse_args()
    config = load_config(args.config)
    if args.model_config:
        import json
        with open(args.model_config) as f:
            config["model_config"] = json.load(f)
    if args.cpu:
        config["device_type"] = "cpu"
    setup_distributed(device_type=config.get("device_type", "cuda"))
    pg = init_process_groups(config.get("device_type", "cuda"), config["mesh_dim"], config["mesh_name"])
    pg.print_mesh_info()
    model = get_strategy("3d", pg, config, checkpoint_path=args.checkpoint, is_staged=True).apply(None)
    n = sum(p.numel() for p in model.parameters())
    print(f"[rank {pg.rank}] stage built: {n / 1e6:.2f} M parameters, coords {pg.get_coordinates_tensor_search(pg.rank)}")
    cleanup_distributed()


if __name__ == "__main__":
    main()
