"""Single-device baseline of the ViT trainer (reference ``examples/train_on_single_gpu.py``)."""
from examples.common import run_vit

if __name__ == "__main__":
    run_vit("dp", "single_config.yaml")
