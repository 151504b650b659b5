#!/bin/bash
mkdir -p gpurun_out
echo "== gemm tests, forced CTA pairs"; QN_GEMM_2CTA=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 --tb=line -x -k "gemm" 2>&1 | tail -4
echo "== all single-GPU tests (default: auto CTA pairs)"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -q -m gpu --timeout 600 --tb=short 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_all.log
echo "== bench gemm"; timeout 200 python scripts/bench_gemm.py 2>&1 | grep "2cta=\|Error\|watchdog" | tee gpurun_out/bench_gemm_2cta.log
echo "== bench"; timeout 600 python bench.py --steps 4 --warmup 3 2>&1 | grep '^{"metric\|Error\|error' | tee gpurun_out/bench_ours.log | cut -c1-420
