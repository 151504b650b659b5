#!/bin/bash
mkdir -p gpurun_out
run() { port=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e "$@" 2>&1 | grep '^{"metric\|Error\|error\|Traceback' | tail -3 | cut -c1-420; }
echo "== tp2 SP fused (default)"; run 29811 --mesh 1,2,1 | tee gpurun_out/bench_tp2_fused.log
echo "== tp2 AR"; run 29812 --mesh 1,2,1 --sequence-parallel 0 | tee gpurun_out/bench_tp2_ar.log
echo "== dp2"; run 29813 | tee gpurun_out/bench_dp2.log
echo "== pp2"; run 29814 --mesh 1,1,2 | tee gpurun_out/bench_pp2.log
echo "== profile tp2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29815 scripts/profile_tp.py 2>&1 | grep -E "qn::|void at|total|nccl|Memcpy|Memset|elementwise" | head -36 | tee gpurun_out/profile_tp2.log
