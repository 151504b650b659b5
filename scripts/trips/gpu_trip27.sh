#!/bin/bash
mkdir -p gpurun_out
echo "== fp8 tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 --tb=short -k "fp8" 2>&1 | tail -15
echo "== fp8 bench"; timeout 300 python scripts/bench_fp8.py 2>&1 | grep "us\|Error" | tee gpurun_out/bench_fp8.log
echo "== all single-GPU tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -q -m gpu --timeout 600 --tb=short 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_all.log
