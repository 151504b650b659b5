#!/bin/bash
# compute-sanitizer targets for the single-GPU kernels (memcheck / racecheck / synccheck).
# usage (on a GPU box): bash scripts/sanitize.sh [memcheck|racecheck|synccheck]
tool=${1:-memcheck}
mkdir -p gpurun_out
compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x \
  -k "layernorm or gelu_dropout_rope or adamw or epilogues or attention_qkvpacked or split_k" 2>&1 | tail -40 | tee gpurun_out/sanitize_$tool.log
