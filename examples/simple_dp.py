"""Pure DP training of the ViT (reference ``examples/simple_dp.py``).

    torchrun --nproc-per-node 2 -m examples.simple_dp --config examples/dp_config.yaml --synthetic [--cpu]
"""
from examples.common import run_vit

if __name__ == "__main__":
    run_vit("dp", "dp_config.yaml")
