#!/bin/bash
# first GPU trip: kernel numerics, smoke, 1-GPU bench (ours + reference)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu" ; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 120 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "== bench ours (small, quick)" ; timeout 600 python bench.py --model gpt2-small --steps 3 --warmup 2 2>&1 | tail -6 | tee gpurun_out/bench_ours_small.log
echo "== bench ours (medium)" ; timeout 900 python bench.py --steps 4 --warmup 3 2>&1 | tail -6 | tee gpurun_out/bench_ours.log
echo "== bench ours, library GEMM (for comparison)" ; QUINTNET_NATIVE_DISABLE=gemm,cross_entropy timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e 2>&1 | tail -4 | tee gpurun_out/bench_ours_nogemm.log
echo "== bench reference" ; timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -6 | tee gpurun_out/bench_ref.log
