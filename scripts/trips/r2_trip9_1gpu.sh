#!/bin/bash
# Round-2 trip 9 (1 GPU): MXFP8 block-scaled GEMM tests, full suite, N=1 step profile.
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip9.log
: > $LOG
step() { name=$1; shift; echo "== $name" | tee -a $LOG; timeout --signal=KILL "$@" >> $LOG 2>&1; echo "   rc=$?" | tee -a $LOG; }
PT="python -m pytest -q -m gpu --timeout 180 -p no:cacheprovider"
step "build (incremental)" 300 python -c "import __graft_entry__ as g; g.build()"
step "mxfp8 tests" 200 $PT tests/test_kernels_gpu.py -k "mxfp8 or fp8"
step "full GPU suite" 700 $PT tests
step "fp8 bench (rowwise vs mxfp8 vs bf16)" 200 python scripts/bench_fp8.py
step "profile N=1 step" 300 python scripts/profile_step.py
tail -c 6000 $LOG
