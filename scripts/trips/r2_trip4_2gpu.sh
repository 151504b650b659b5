#!/bin/bash
# Round-2 trip 4 (2 GPUs): the rewritten fused TP kernels over real NVLink + everything multi-GPU written so far.
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip4.log
: > $LOG
step() { name=$1; shift; echo "== $name" | tee -a $LOG; timeout --signal=KILL "$@" >> $LOG 2>&1; echo "   rc=$?" | tee -a $LOG; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m >> $LOG 2>&1
step "comm tests (2 GPUs)" 300 python -m pytest tests/test_comm_gpu.py -q -m gpu -x --timeout 120 -p no:cacheprovider
step "tp microbench pump/lsu x pairs" 200 $TR --master-port 29860 scripts/bench_tp_kernels.py
step "parity tp2 / dp2 / pp2 vs 1 GPU" 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x --timeout 500 -p no:cacheprovider
run() { name=$1; port=$2; shift 2; echo "== bench $name" | tee -a $LOG; timeout --signal=KILL 240 $TR --master-port $port bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e "$@" 2>&1 | grep '^{"metric\|Error\|error\|Traceback' | tail -3 | cut -c1-700 | tee -a $LOG; }
run "tp2 pump pairs (default)" 29841 --mesh 1,2,1
QN_TP_AG_MODE=lsu run "tp2 lsu pairs" 29842 --mesh 1,2,1
QN_TP_2CTA=0 run "tp2 pump single-cta" 29847 --mesh 1,2,1
run "pp2 plain" 29843 --mesh 1,1,2
run "pp2 interleaved V=2" 29844 --mesh 1,1,2 --virtual-stages 2
run "dp2 zero1" 29846
echo "== bench reference bf16 dp2" | tee -a $LOG; timeout --signal=KILL 300 $TR --master-port 29848 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference 2>&1 | grep '^{"' | cut -c1-500 | tee -a $LOG
tail -c 6000 $LOG
