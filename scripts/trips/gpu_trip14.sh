#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 180 --tb=short 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench ours (medium)"; timeout 900 python bench.py --steps 4 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours.log
echo "== profile"; timeout 600 python scripts/profile_step.py 2>&1 | grep -E "qn::|void at|total device|cudnn|nccl|Memcpy|Memset" | head -24 | tee gpurun_out/profile_step.log
