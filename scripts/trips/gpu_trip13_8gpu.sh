#!/bin/bash
mkdir -p gpurun_out
run() { n=$1; port=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" 2>&1 | grep '^{"metric\|Error\|error\|Traceback\|WARNING' | tail -4 | cut -c1-2000; }
echo "== ours N=8"; run 8 29901 --steps 4 --warmup 3 | tee gpurun_out/bench_ours_8.log
echo "== ours N=4"; run 4 29902 --steps 4 --warmup 3 | tee gpurun_out/bench_ours_4.log
echo "== ours N=2"; run 2 29903 --steps 4 --warmup 3 | tee gpurun_out/bench_ours_2.log
