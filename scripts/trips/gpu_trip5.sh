#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 180 --tb=line 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "== kernel times"; timeout 600 python scripts/ncu_targets.py all --time 2>&1 | tail -40 | tee gpurun_out/kernel_times.log
echo "== bench ours (medium)"; timeout 900 python bench.py --steps 4 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours.log
echo "== profile"; timeout 600 python scripts/profile_step.py 2>&1 | grep -E "qn::|void at|total device|cudnn|nccl" | head -40 | tee gpurun_out/profile_step.log
