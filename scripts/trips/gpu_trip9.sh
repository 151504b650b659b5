#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 180 --tb=line 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "== kernel times"; timeout 600 python scripts/ncu_targets.py all --time 2>&1 | tail -60 | tee gpurun_out/kernel_times.log
echo "== bench ours (medium)"; timeout 900 python bench.py --steps 4 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours.log
echo "== ncu gelu gemm"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 40 -c 1 -o gpurun_out/prof_gemm_gelu -f python scripts/ncu_gelu.py 2>&1 | tail -2
