#!/bin/bash
mkdir -p gpurun_out
echo "== LN tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 --tb=short -k "layernorm" 2>&1 | tail -15
echo "== LN bench"; timeout 300 python scripts/bench_ln.py 2>&1 | grep "us \|Error" | tee gpurun_out/bench_ln.log
echo "== bench"; timeout 600 python bench.py --steps 4 --warmup 3 2>&1 | grep '^{"metric\|Error\|error' | tee gpurun_out/bench_ours.log | cut -c1-400
