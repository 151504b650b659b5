#!/bin/bash
mkdir -p gpurun_out
run() { n=$1; port=$2; shift 2
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" 2>&1 | grep '^{"metric\|Error\|error\|Traceback\|WARNING' | tail -4 | cut -c1-2000; }
echo "== ours N=4 default"; run 4 29901 --steps 4 --warmup 3 | tee gpurun_out/bench_ours_4_v3.log
echo "== ours N=4 tp2 pp2"; run 4 29902 --steps 4 --warmup 3 --mesh 1,2,2 | tee gpurun_out/bench_ours_4_tp2pp2.log
echo "== vit ours N=4"; run 4 29903 --steps 10 --warmup 3 --workload vit | tee gpurun_out/bench_vit_ours_4.log
echo "== vit ref N=4"; run 4 29904 --steps 10 --warmup 3 --workload vit --impl reference | tee gpurun_out/bench_vit_ref_4.log
