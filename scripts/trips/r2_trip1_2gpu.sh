#!/bin/bash
# Round-2 trip 1 (2 GPUs): validate the shipped binary + every opt-in path written at the end of round 1.
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip1.log
: > $LOG
step() { name=$1; shift; echo "== $name" | tee -a $LOG; timeout --signal=KILL "$@" >> $LOG 2>&1; echo "   rc=$?" | tee -a $LOG; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m >> $LOG 2>&1
step "comm tests default (final binary)" 200 python -m pytest tests/test_comm_gpu.py -q -m gpu -x --timeout 120
QN_TP_2CTA=1 step "tp fused tests, CTA pairs" 150 python -m pytest tests/test_comm_gpu.py -q -m gpu -x --timeout 90 -k "tp_fused or sequence_parallel"
step "tp microbench default" 120 $TR --master-port 29860 scripts/bench_tp_kernels.py
QN_TP_2CTA=1 step "tp microbench CTA pairs" 120 $TR --master-port 29861 scripts/bench_tp_kernels.py
step "copy-engine overlap experiment" 120 $TR --master-port 29871 scripts/exp_copy_engine_overlap.py
QN_TP_AG_MODE=ce step "tp fused tests, copy-engine AG" 150 python -m pytest tests/test_comm_gpu.py -q -m gpu -x --timeout 90 -k "tp_fused or sequence_parallel"
QN_TP_AG_MODE=ce step "tp microbench copy-engine AG" 120 $TR --master-port 29862 scripts/bench_tp_kernels.py
run() { name=$1; port=$2; shift 2; echo "== bench $name" | tee -a $LOG; timeout --signal=KILL 240 $TR --master-port $port bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e "$@" 2>&1 | grep '^{"metric\|Error\|error\|Traceback' | tail -3 | cut -c1-400 | tee -a $LOG; }
run "tp2 default" 29841 --mesh 1,2,1
QN_TP_2CTA=1 run "tp2 CTA pairs" 29842 --mesh 1,2,1
QN_TP_AG_MODE=ce run "tp2 copy-engine AG" 29845 --mesh 1,2,1
run "pp2 plain" 29843 --mesh 1,1,2
run "pp2 interleaved V=2" 29844 --mesh 1,1,2 --virtual-stages 2
run "dp2 zero1" 29846
QN_FP8_2CTA=1 step "fp8 CTA pairs tests" 90 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 60 -k fp8
QN_FP8_2CTA=1 step "fp8 bench" 90 python scripts/bench_fp8.py
QN_RUN_UNVALIDATED=1 step "unvalidated GPU tests" 150 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 90 -k "activation_checkpointing or wgrad_accumulate"
echo "== N=1 bench fused wgrad accum" | tee -a $LOG; QUINTNET_FUSE_WGRAD_ACCUM=1 timeout --signal=KILL 240 python bench.py --steps 3 --warmup 3 --no-e2e 2>&1 | grep '^{"metric' | cut -c1-400 | tee -a $LOG
echo "== N=1 bench default" | tee -a $LOG; timeout --signal=KILL 240 python bench.py --steps 3 --warmup 3 --no-e2e 2>&1 | grep '^{"metric' | cut -c1-400 | tee -a $LOG
tail -c 6000 $LOG
