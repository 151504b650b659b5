#!/bin/bash
mkdir -p gpurun_out
echo "== nvlink push bench"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29871 scripts/bench_nvlink_push.py 2>&1 | grep "GB/s\|Error\|error\|16 MiB" | tee gpurun_out/nvlink_push.log
echo "== tp kernel microbench"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29861 scripts/bench_tp_kernels.py 2>&1 | grep "dbg=\|Error\|error" | tee gpurun_out/tp_kernel_times.log
