#!/bin/bash
mkdir -p gpurun_out
echo "== gemm tests, forced CTA pairs"; QN_GEMM_2CTA=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 --tb=line -x -k "gemm_tn or gemm_layouts or gemm_epilogues" 2>&1 | tail -8
echo "== bench 1cta"; QN_GEMM_2CTA=0 timeout 200 python scripts/bench_gemm.py 2>&1 | grep "2cta=\|Error" | tee gpurun_out/bench_gemm_1cta.log
echo "== bench 2cta"; QN_GEMM_2CTA=1 timeout 200 python scripts/bench_gemm.py 2>&1 | grep "2cta=\|Error\|watchdog" | tee gpurun_out/bench_gemm_2cta.log
