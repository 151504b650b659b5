#!/bin/bash
mkdir -p gpurun_out
echo "== pytest attention"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 --tb=short -k "attention or gpt2" 2>&1 | tail -8 | tee gpurun_out/pytest_attn.log
echo "== kernel times"; timeout 600 python scripts/ncu_targets.py attn --time 2>&1 | tail -8 | tee gpurun_out/kernel_times_attn.log
echo "== bench ours (medium)"; timeout 900 python bench.py --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-330 | tee gpurun_out/bench_ours.log
