#!/bin/bash
# Round-2 trip 6 (2 GPUs): done-flag fused TP kernels over real NVLink, ViT tp2/pp2 with CUDA graphs + one-shot AR.
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip6.log
: > $LOG
step() { name=$1; shift; echo "== $name" | tee -a $LOG; timeout --signal=KILL "$@" >> $LOG 2>&1; echo "   rc=$?" | tee -a $LOG; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
step "build (incremental)" 300 python -c "import __graft_entry__ as g; g.build()"
step "comm tests (2 GPUs)" 300 python -m pytest tests/test_comm_gpu.py -q -m gpu -x --timeout 120 -p no:cacheprovider
step "nvlink push vs pull" 200 $TR --master-port 29859 scripts/bench_nvlink_rw.py
step "tp microbench + attribution" 300 $TR --master-port 29860 scripts/bench_tp_kernels.py
step "parity tp2 / dp2 / pp2 vs 1 GPU" 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x --timeout 500 -p no:cacheprovider
run() { name=$1; port=$2; shift 2; echo "== bench $name" | tee -a $LOG; timeout --signal=KILL 240 $TR --master-port $port bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e "$@" 2>&1 | grep '^{"metric\|Error\|error\|Traceback' | tail -3 | cut -c1-900 | tee -a $LOG; }
run "tp2 (default)" 29841 --mesh 1,2,1
run "pp2 (direction channels + pre-posted recv)" 29843 --mesh 1,1,2
run "dp2 zero1 (NVLS phases)" 29846
vit() { name=$1; port=$2; shift 2; echo "== vit $name" | tee -a $LOG; timeout --signal=KILL 200 $TR --master-port $port bench.py --workload vit --gpus 2 --steps 30 --warmup 5 "$@" 2>&1 | grep '^{"metric\|rror\|graph' | tail -4 | cut -c1-700 | tee -a $LOG; }
vit "ours pp2 graphs" 29851 --mesh 1,1,2
QN_VIT_GRAPHS=0 vit "ours pp2 eager" 29852 --mesh 1,1,2
vit "ours tp2" 29853 --mesh 1,2,1
vit "reference pp2" 29854 --mesh 1,1,2 --impl reference
vit "reference tp2" 29855 --mesh 1,2,1 --impl reference
step "profile tp2 step" 200 $TR --master-port 29870 scripts/profile_tp.py
tail -c 6000 $LOG
