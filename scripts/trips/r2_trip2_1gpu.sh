#!/bin/bash
# Round-2 trip 2 (1 GPU): loop-back validation of the rewritten fused TP kernels + the single-GPU suite + bench.
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip2.log
: > $LOG
step() { name=$1; shift; echo "== $name" | tee -a $LOG; timeout --signal=KILL "$@" >> $LOG 2>&1; echo "   rc=$?" | tee -a $LOG; }
PT="python -m pytest -q -m gpu -x --timeout 120 -p no:cacheprovider"
step "loopback AG single-cta lsu"   200 $PT tests/test_tp_loopback_gpu.py -k "ag_gemm and single-cta and lsu"
step "loopback AG single-cta pump"  200 $PT tests/test_tp_loopback_gpu.py -k "ag_gemm and single-cta and pump"
step "loopback AG cta-pairs pump"   200 $PT tests/test_tp_loopback_gpu.py -k "ag_gemm and cta-pairs and pump"
step "loopback AG cta-pairs lsu"    200 $PT tests/test_tp_loopback_gpu.py -k "ag_gemm and cta-pairs and lsu"
step "loopback RS single-cta"       200 $PT tests/test_tp_loopback_gpu.py -k "gemm_rs and single-cta"
step "loopback RS cta-pairs"        200 $PT tests/test_tp_loopback_gpu.py -k "gemm_rs and cta-pairs"
step "kernel + model GPU suite"     400 $PT tests/test_kernels_gpu.py tests/test_models_gpu.py
echo "== N=1 bench" | tee -a $LOG; timeout --signal=KILL 300 python bench.py --steps 3 --warmup 3 2>&1 | grep '^{"metric' | cut -c1-1500 | tee -a $LOG
echo "== N=1 reference both precisions" | tee -a $LOG; timeout --signal=KILL 400 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | grep '^{"\|Error\|error' | cut -c1-1500 | tee -a $LOG
tail -c 5000 $LOG
