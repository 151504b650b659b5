#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 180 --tb=line 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "== debug"; timeout 300 python scripts/debug_smoke.py 2>&1 | tail -45 | tee gpurun_out/debug_smoke.log
echo "== kernel times"; timeout 600 python scripts/ncu_targets.py all --time 2>&1 | tail -40 | tee gpurun_out/kernel_times.log
echo "== ncu attn fwd"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -c 1 -o gpurun_out/prof_attn_fwd -f python scripts/ncu_targets.py attn 2>&1 | tail -3
echo "== ncu attn bwd"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -c 1 -o gpurun_out/prof_attn_bwd -f python scripts/ncu_targets.py attn 2>&1 | tail -3
echo "== ncu gemm"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -c 1 -o gpurun_out/prof_gemm_tn -f python scripts/ncu_targets.py gemm 2>&1 | tail -3
echo "== bench ours (medium)"; timeout 900 python bench.py --steps 4 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours.log
ls -la gpurun_out/
