#!/bin/bash
mkdir -p gpurun_out
echo "== profile tp2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29835 scripts/profile_tp.py 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" | tail -42 | cut -c1-200 | tee gpurun_out/profile_tp2.log
