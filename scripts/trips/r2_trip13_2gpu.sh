#!/bin/bash
# Round-2 trip 13 (2 GPUs, short): BASELINE config #3 (GPT-2 small, TP = 2) both arms, tp2 parity and comm tests with the final binary.
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip13.log
: > $LOG
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
step() { name=$1; t=$2; shift 2; echo "== $name" >> $LOG; timeout --signal=TERM -k 10 $t "$@" >> $LOG 2>&1; echo "   rc=$?" >> $LOG; }
step "build (incremental)" 200 python -c "import __graft_entry__ as g; g.build()"
step "gpt2-small tp2 ours" 150 $TR --master-port 29871 bench.py --gpus 2 --mesh 1,2,1 --model gpt2-small --steps 5 --warmup 3
step "gpt2-small tp2 reference (fp32 + bf16 autocast)" 200 $TR --master-port 29872 bench.py --impl reference --gpus 2 --mesh 1,2,1 --model gpt2-small --steps 3 --warmup 1
step "tp2 parity vs 1 GPU" 200 python -m pytest tests/test_parity_gpu.py -q -m gpu -x --timeout 180 -p no:cacheprovider -k "tp2"
step "comm tests (2 GPUs)" 200 python -m pytest tests/test_comm_gpu.py -q -m gpu -x --timeout 120 -p no:cacheprovider
grep '^{"metric\|^==\|rc=\|passed\|failed' $LOG | cut -c1-400
