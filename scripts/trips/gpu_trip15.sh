#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all single-GPU)"; timeout 1200 python -m pytest tests -q -m gpu --timeout 300 --tb=short -x 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_all.log
