"""Pure TP training of the ViT (reference ``examples/simple_tp.py``).

    torchrun --nproc-per-node 2 -m examples.simple_tp --config examples/tp_config.yaml --synthetic [--cpu]
"""
from examples.common import run_vit

if __name__ == "__main__":
    run_vit("tp", "tp_config.yaml")
