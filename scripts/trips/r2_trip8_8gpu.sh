#!/bin/bash
# Round-2 trip 8 (8 GPUs, charged 8x: keep it short).  Headline 2x2x2 bench (both arms, iso-precision), N=4 bench, DP=8
# all-reduce sweep vs NCCL, ViT 2x2x2 (both arms), 2x2x2-vs-1-GPU loss parity, world-4 in-kernel collectives.
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip8.log
: > $LOG
step() { name=$1; shift; echo "== $name" | tee -a $LOG; timeout --signal=KILL "$@" >> $LOG 2>&1; echo "   rc=$?" | tee -a $LOG; }
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
step "build (incremental)" 300 python -c "import __graft_entry__ as g; g.build()"
echo "== ours N=8 (2x2x2)" | tee -a $LOG; timeout --signal=KILL 200 $TR8 --master-port 29901 bench.py --gpus 8 --steps 5 --warmup 3 2>&1 | grep '^{"metric\|rror\|Traceback' | tail -3 | cut -c1-2500 | tee -a $LOG | tee gpurun_out/r2_bench_ours_8.json
echo "== reference N=8 (fp32 + bf16 autocast)" | tee -a $LOG; timeout --signal=KILL 300 $TR8 --master-port 29902 bench.py --gpus 8 --steps 3 --warmup 1 --impl reference 2>&1 | grep '^{"' | cut -c1-1800 | tee -a $LOG | tee gpurun_out/r2_bench_ref_8.json
echo "== ours N=4 (dp2 x pp2)" | tee -a $LOG; CUDA_VISIBLE_DEVICES=0,1,2,3 timeout --signal=KILL 200 $TR4 --master-port 29903 bench.py --gpus 4 --steps 5 --warmup 3 --no-e2e 2>&1 | grep '^{"metric\|rror\|Traceback' | tail -3 | cut -c1-1500 | tee -a $LOG | tee gpurun_out/r2_bench_ours_4.json
echo "== ours N=8 dp4 x pp2 (no TP, comparison)" | tee -a $LOG; timeout --signal=KILL 200 $TR8 --master-port 29904 bench.py --gpus 8 --mesh 4,1,2 --steps 4 --warmup 3 --no-e2e 2>&1 | grep '^{"metric\|rror\|Traceback' | tail -3 | cut -c1-900 | tee -a $LOG
step "all-reduce sweep DP=8 (K1 vs NCCL, 1 KB - 1 GB)" 240 $TR8 --master-port 29905 scripts/bench_allreduce_sweep.py
echo "== ViT 2x2x2 ours (graphs)" | tee -a $LOG; timeout --signal=KILL 200 $TR8 --master-port 29906 bench.py --workload vit --gpus 8 --steps 30 --warmup 5 2>&1 | grep '^{"metric\|rror\|graph' | tail -3 | cut -c1-900 | tee -a $LOG | tee gpurun_out/r2_bench_vit_ours_8.json
echo "== ViT 2x2x2 ours (eager)" | tee -a $LOG; QN_VIT_GRAPHS=0 timeout --signal=KILL 200 $TR8 --master-port 29907 bench.py --workload vit --gpus 8 --steps 30 --warmup 5 2>&1 | grep '^{"metric\|rror' | tail -2 | cut -c1-600 | tee -a $LOG
echo "== ViT 2x2x2 reference" | tee -a $LOG; timeout --signal=KILL 200 $TR8 --master-port 29908 bench.py --workload vit --gpus 8 --steps 30 --warmup 5 --impl reference 2>&1 | grep '^{"metric\|rror' | tail -2 | cut -c1-900 | tee -a $LOG | tee gpurun_out/r2_bench_vit_ref_8.json
step "parity 3-D 2x2x2 vs 1 GPU" 400 python -m pytest tests/test_parity_gpu.py -q -m gpu -x --timeout 380 -p no:cacheprovider -k "3d"
step "in-kernel collectives world 4" 300 python -m pytest tests/test_comm_gpu.py -q -m gpu --timeout 200 -p no:cacheprovider -k "world4"
tail -c 8000 $LOG
