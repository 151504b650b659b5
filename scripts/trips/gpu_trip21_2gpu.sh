#!/bin/bash
mkdir -p gpurun_out
echo "== comm tests"; timeout 600 python -m pytest tests/test_comm_gpu.py -q -m gpu --timeout 300 --tb=short -s 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/pytest_comm.log
echo "== tp2 profile"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29851 scripts/profile_tp.py 2>&1 | grep -v "Warning\|warn" | tee gpurun_out/profile_tp2.log | head -60
