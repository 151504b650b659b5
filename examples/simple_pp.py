"""Pure PP training of the ViT (reference ``examples/simple_pp.py``).

    torchrun --nproc-per-node 2 -m examples.simple_pp --config examples/pp_config.yaml --synthetic [--cpu]
"""
from examples.common import run_vit

if __name__ == "__main__":
    run_vit("pp", "pp_config.yaml")
