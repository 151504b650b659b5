#!/bin/bash
# Round-2 hardware validation of everything that was written but could not be run in round 1.
# EVERY step carries its own short `timeout` (a hung cluster barrier has no device-side watchdog); run under
#   gpurun --gpus 2 --timeout 420 -- 'bash scripts/validate_round2.sh'
# and never with a gpurun limit close to the remaining budget.
mkdir -p gpurun_out
step() { name=$1; shift; echo "== $name"; timeout --signal=KILL "$@" 2>&1 | tail -6; echo "   rc=$?"; }
# 1. fp8 on CTA pairs (1 GPU)
QN_FP8_2CTA=1 step "fp8 CTA pairs" 90 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 60 -k fp8
QN_FP8_2CTA=1 step "fp8 bench" 90 python scripts/bench_fp8.py
# 2. fused TP kernels on CTA pairs (2 GPUs)
QN_TP_2CTA=1 step "tp fused tests, CTA pairs" 120 python -m pytest tests/test_comm_gpu.py -q -m gpu -x --timeout 90 -k "tp_fused or sequence_parallel"
QN_TP_2CTA=1 step "tp kernel microbench, CTA pairs" 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29861 scripts/bench_tp_kernels.py
# 3. tp2 end to end
run() { port=$1; shift; timeout --signal=KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e "$@" 2>&1 | grep '^{"metric\|Error\|error\|Traceback' | tail -3 | cut -c1-330; }
echo "== tp2 SP fused, single-CTA fused kernels"; run 29841 --mesh 1,2,1
echo "== tp2 SP fused, CTA-pair fused kernels"; QN_TP_2CTA=1 run 29842 --mesh 1,2,1
# 4. interleaved 1F1B (virtual pipeline stages) vs plain 1F1B at pp2 (CPU-validated; needs its first NCCL run)
echo "== pp2 plain 1F1B"; run 29843 --mesh 1,1,2
echo "== pp2 interleaved V=2"; run 29844 --mesh 1,1,2 --virtual-stages 2
# 5. design input for the fused AG kernels: does NVLink traffic alone slow a concurrent GEMM?
step "copy-engine overlap experiment" 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29871 scripts/exp_copy_engine_overlap.py
# 6. GPU tests written without hardware access
QN_RUN_UNVALIDATED=1 step "unvalidated GPU tests" 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 90 -k "activation_checkpointing or wgrad_accumulate"
# 7. all-gather by the copy engine instead of push warps (decides NEXT.md section 2, option 1)
QN_TP_AG_MODE=ce step "tp fused tests, copy-engine AG" 120 python -m pytest tests/test_comm_gpu.py -q -m gpu -x --timeout 90 -k "tp_fused or sequence_parallel"
echo "== tp2 SP fused, copy-engine AG"; QN_TP_AG_MODE=ce run 29845 --mesh 1,2,1
# 8. fused weight-gradient accumulation end to end (N = 1 bench with and without)
echo "== N=1 bench, fused wgrad accumulation"; QUINTNET_FUSE_WGRAD_ACCUM=1 timeout --signal=KILL 240 python bench.py --steps 3 --warmup 3 --no-e2e 2>&1 | grep '^{"metric' | cut -c1-330
