#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== debug smoke"; timeout 300 python scripts/debug_smoke.py 2>&1 | tail -60 | tee gpurun_out/debug_smoke.log
echo "== profile"; timeout 600 python scripts/profile_step.py 2>&1 | tail -50 | tee gpurun_out/profile_step.log
