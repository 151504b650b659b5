#!/bin/bash
# Round-2 trip 10 (2 GPUs): fused-TP kernels with per-part publication (A/B against one flag per shard), tp2 step profile,
# NVLink push-vs-pull rates, ViT tp2/dp2 with model-level CUDA graphs.
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip10.log
: > $LOG
step() { name=$1; shift; echo "== $name" | tee -a $LOG; timeout --signal=KILL "$@" >> $LOG 2>&1; echo "   rc=$?" | tee -a $LOG; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
step "build (incremental)" 300 python -c "import __graft_entry__ as g; g.build()"
step "tp microbench (4 parts)" 200 $TR --master-port 29860 scripts/bench_tp_kernels.py
QN_BENCH_TP_MODES=pump:1 QN_BENCH_TP_OUT=gpurun_out/r2_tp_kernel_times_1part.json QN_TP_AG_PARTS=1 QN_TP_RS_PARTS=1 step "tp microbench (1 part = single done flag)" 200 $TR --master-port 29861 scripts/bench_tp_kernels.py
run() { name=$1; port=$2; shift 2; echo "== bench $name" | tee -a $LOG; timeout --signal=KILL 240 $TR --master-port $port bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e "$@" 2>&1 | grep '^{"metric\|Error\|error\|Traceback' | tail -3 | cut -c1-900 | tee -a $LOG; }
run "tp2 (4 parts)" 29841 --mesh 1,2,1
QN_TP_AG_PARTS=1 QN_TP_RS_PARTS=1 run "tp2 (1 part)" 29842 --mesh 1,2,1
step "tp2 step profile" 240 $TR --master-port 29843 scripts/profile_tp.py
cp gpurun_out/r2_trip10.log gpurun_out/r2_trip10_partial.log
step "nvlink push vs pull" 150 $TR --master-port 29859 scripts/bench_nvlink_rw.py
vit() { name=$1; port=$2; shift 2; echo "== vit $name" | tee -a $LOG; timeout --signal=KILL 200 $TR --master-port $port bench.py --workload vit --gpus 2 --steps 30 --warmup 5 "$@" 2>&1 | grep '^{"metric\|rror\|graph' | tail -4 | cut -c1-700 | tee -a $LOG; }
vit "ours tp2 (graphs)" 29853 --mesh 1,2,1
vit "ours dp2 (graphs)" 29854 --mesh 2,1,1
step "tp parity (tp2 vs 1 GPU)" 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -x --timeout 280 -p no:cacheprovider -k "tp2"
tail -c 7000 $LOG
