#!/bin/bash
# Round-2 trip 3 (1 GPU): full single-GPU suite with fp32 main grads, smoke, N=1 bench, kernel timings, ncu full set,
# compute-sanitizer on the loop-back fused-TP tests and a kernel subset.
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip3.log
: > $LOG
step() { name=$1; shift; echo "== $name" | tee -a $LOG; timeout --signal=KILL "$@" >> $LOG 2>&1; echo "   rc=$?" | tee -a $LOG; }
PT="python -m pytest -q -m gpu --timeout 180 -p no:cacheprovider"
step "full GPU suite" 600 $PT tests
step "smoke" 120 python -c "import __graft_entry__ as g; g.smoke()"
echo "== N=1 bench" | tee -a $LOG; timeout --signal=KILL 300 python bench.py --steps 5 --warmup 3 2>&1 | grep '^{"metric' | cut -c1-1800 | tee -a $LOG | tee gpurun_out/r2_bench_ours_1.json
echo "== N=1 bench, bf16 grads (main_grad_dtype=param)" | tee -a $LOG; QN_BENCH_MAIN_GRAD=param timeout --signal=KILL 300 python bench.py --steps 3 --warmup 3 --no-e2e 2>&1 | grep '^{"metric' | cut -c1-600 | tee -a $LOG
step "kernel timings" 200 python scripts/ncu_r2_targets.py --time
step "ncu full set" 500 ncu --set full --clock-control none --import-source on --profile-from-start off --nvtx -f -o gpurun_out/r2_prof python scripts/ncu_r2_targets.py
step "sanitizer memcheck loopback" 240 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_tp_loopback_gpu.py -q -m gpu -x -p no:cacheprovider -k "2048 and cta-pairs and (pump or gemm_rs)"
step "sanitizer synccheck loopback" 240 compute-sanitizer --tool synccheck --print-limit 10 python -m pytest tests/test_tp_loopback_gpu.py -q -m gpu -x -p no:cacheprovider -k "2048 and cta-pairs and (pump or gemm_rs)"
step "sanitizer memcheck kernels" 300 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -p no:cacheprovider -k "layernorm or adamw or epilogues or attention_qkvpacked or split_k or main_grad"
tail -c 6000 $LOG
