#!/bin/bash
mkdir -p gpurun_out
echo "== comm tests"; timeout 600 python -m pytest tests/test_comm_gpu.py -q -m gpu --timeout 300 --tb=short 2>&1 | tail -12 | tee gpurun_out/pytest_comm.log
run() { port=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e "$@" 2>&1 | grep '^{"metric\|Error\|error\|Traceback' | tail -3 | cut -c1-420; }
echo "== tp2 SP fused (default)"; run 29821 --mesh 1,2,1 | tee gpurun_out/bench_tp2_fused.log
echo "== profile tp2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29825 scripts/profile_tp.py 2>&1 | grep -E "qn::|void at|total|nccl|Memcpy|Memset|elementwise" | head -16 | tee gpurun_out/profile_tp2.log
