#!/bin/bash
mkdir -p gpurun_out
echo "== comm tests"; timeout 600 python -m pytest tests/test_comm_gpu.py -q -m gpu --timeout 300 --tb=short -x 2>&1 | tail -30 | tee gpurun_out/pytest_comm.log
run() { port=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --mesh 1,2,1 --steps 3 --warmup 2 --no-e2e "$@" 2>&1 | grep '^{"metric\|Error\|error' | tail -3; }
echo "== tp2 plain (AR)"; run 29801 | tee gpurun_out/bench_tp2_ar.log
echo "== tp2 SP unfused"; run 29802 --sequence-parallel 1 | tee gpurun_out/bench_tp2_sp_unfused.log
echo "== tp2 SP fused"; QUINTNET_NATIVE_ENABLE=tp_fused run 29803 --sequence-parallel 1 | tee gpurun_out/bench_tp2_sp_fused.log
