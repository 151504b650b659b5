#!/bin/bash
mkdir -p gpurun_out
echo "== tp fused tests (comm CTAs 16)"; QN_TP_COMM_CTAS=16 timeout 600 python -m pytest tests/test_comm_gpu.py -q -m gpu --timeout 300 --tb=short -k "tp_fused or sequence_parallel" 2>&1 | tail -4

echo "== tp kernel microbench"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29861 scripts/bench_tp_kernels.py 2>&1 | grep "dbg=\|Error\|error" | tee gpurun_out/tp_kernel_times.log
run() { port=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e "$@" 2>&1 | grep '^{"metric\|Error\|error\|Traceback' | tail -3 | cut -c1-330; }
echo "== tp2 SP fused comm16"; QN_TP_COMM_CTAS=16 run 29841 --mesh 1,2,1 | tee gpurun_out/bench_tp2_fused_comm16.log
