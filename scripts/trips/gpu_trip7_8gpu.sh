#!/bin/bash
mkdir -p gpurun_out
run() { # n port extra...
  n=$1; port=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" 2>&1 | grep -v "Training Epoch\|OMP_NUM\|\*\*\*\*" | tail -3
}
echo "== ours N=8"; run 8 29701 --steps 4 --warmup 3 | tee gpurun_out/bench_ours_8.log
echo "== ours N=4"; run 4 29702 --steps 4 --warmup 3 | tee gpurun_out/bench_ours_4.log
echo "== ours N=8 sequence-parallel (unfused collectives)"; run 8 29703 --steps 3 --warmup 2 --sequence-parallel 1 --no-e2e | tee gpurun_out/bench_ours_8_sp.log
echo "== ours N=8 dp4 pp2"; run 8 29704 --steps 3 --warmup 2 --mesh 4,1,2 --no-e2e | tee gpurun_out/bench_ours_8_dp4pp2.log
echo "== reference N=8"; run 8 29705 --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_ref_8.log
echo "== reference N=4"; run 4 29706 --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_ref_4.log
