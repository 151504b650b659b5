#!/bin/bash
# Round-2 trip 7 (1 GPU): warp-convergent MMA/TMA issue (GEMM + attention), LN-bwd bias-gradient fusion, smoke.
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip7.log
: > $LOG
step() { name=$1; shift; echo "== $name" | tee -a $LOG; timeout --signal=KILL "$@" >> $LOG 2>&1; echo "   rc=$?" | tee -a $LOG; }
PT="python -m pytest -q -m gpu --timeout 180 -p no:cacheprovider"
step "build (incremental)" 300 python -c "import __graft_entry__ as g; g.build()"
step "full GPU suite" 700 $PT tests
step "smoke" 120 python -c "import __graft_entry__ as g; g.smoke()"
step "sustained GEMM ours vs cuBLAS" 200 python scripts/bench_gemm_sustained.py
step "kernel timings" 200 python scripts/ncu_r2_targets.py --time
echo "== N=1 bench" | tee -a $LOG; timeout --signal=KILL 300 python bench.py --steps 5 --warmup 3 2>&1 | grep '^{"metric' | cut -c1-1800 | tee -a $LOG | tee gpurun_out/r2_bench_ours_1_trip7.json
step "ncu full set" 500 ncu --set full --clock-control none --import-source on --profile-from-start off --nvtx -f -o gpurun_out/r2_prof_trip7 python scripts/ncu_r2_targets.py
tail -c 5000 $LOG
