#!/bin/bash
# Round-2 trip 5 (1 GPU): done-flag fused TP kernels (loop-back), new kernels/tests, sustained GEMM vs cuBLAS with the
# sleeping barrier waits, kernel timings, N=1 bench, ViT bench.
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip5.log
: > $LOG
step() { name=$1; shift; echo "== $name" | tee -a $LOG; timeout --signal=KILL "$@" >> $LOG 2>&1; echo "   rc=$?" | tee -a $LOG; }
PT="python -m pytest -q -m gpu --timeout 180 -p no:cacheprovider"
step "build (incremental)" 300 python -c "import __graft_entry__ as g; g.build()"
step "loopback fused TP" 400 $PT -x tests/test_tp_loopback_gpu.py
step "full GPU suite" 600 $PT tests --deselect tests/test_tp_loopback_gpu.py
step "smoke" 120 python -c "import __graft_entry__ as g; g.smoke()"
step "sustained GEMM ours vs cuBLAS" 200 python scripts/bench_gemm_sustained.py
step "kernel timings" 200 python scripts/ncu_r2_targets.py --time
echo "== N=1 bench" | tee -a $LOG; timeout --signal=KILL 300 python bench.py --steps 5 --warmup 3 2>&1 | grep '^{"metric' | cut -c1-1800 | tee -a $LOG | tee gpurun_out/r2_bench_ours_1_trip5.json
echo "== ViT N=1 ours" | tee -a $LOG; timeout --signal=KILL 200 python bench.py --workload vit --steps 20 --warmup 5 2>&1 | grep '^{"metric\|rror' | cut -c1-900 | tee -a $LOG
echo "== ViT N=1 reference" | tee -a $LOG; timeout --signal=KILL 200 python bench.py --workload vit --impl reference --steps 20 --warmup 5 2>&1 | grep '^{"metric\|rror' | cut -c1-900 | tee -a $LOG
tail -c 7000 $LOG
