#!/bin/bash
# Round-2 final single-GPU validation of the frozen tree: build, full GPU suite, smoke, bench N = 1.
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip14.log
: > $LOG
step() { name=$1; t=$2; shift 2; echo "== $name" >> $LOG; timeout --signal=TERM -k 10 $t "$@" >> $LOG 2>&1; echo "   rc=$?" >> $LOG; }
step "build (incremental)" 200 python -c "import __graft_entry__ as g; g.build()"
step "pytest -m gpu" 400 python -m pytest tests -q -m gpu -x --timeout 200 -p no:cacheprovider
step "smoke" 150 python -c "import __graft_entry__ as g; g.smoke()"
step "bench N=1" 200 python bench.py --gpus 1 --steps 5 --warmup 3
grep '^{"metric' $LOG | tail -1 > gpurun_out/r2_bench_ours_1_final.json
grep '^==\|rc=\|passed\|failed\|smoke\|^{"metric' $LOG | cut -c1-600
