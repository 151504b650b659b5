#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8 | tee gpurun_out/topo.txt
echo "== comm tests"; timeout 600 python -m pytest tests/test_comm_gpu.py -q -m gpu --timeout 300 --tb=short -s 2>&1 | tail -40 | tee gpurun_out/pytest_comm.log
echo "== bench ours N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 4 --warmup 3 2>&1 | tail -4 | tee gpurun_out/bench_ours_2.log
echo "== bench ours N=2 (NCCL collectives)"; QUINTNET_NATIVE_DISABLE=allreduce,zero1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 4 --warmup 3 --no-e2e 2>&1 | tail -2 | tee gpurun_out/bench_ours_2_nccl.log
echo "== bench ours N=2 tp2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus 2 --mesh 1,2,1 --steps 3 --warmup 2 --no-e2e 2>&1 | tail -2 | tee gpurun_out/bench_ours_2_tp.log
echo "== bench ours N=2 pp2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29604 bench.py --gpus 2 --mesh 1,1,2 --steps 3 --warmup 2 --no-e2e 2>&1 | tail -2 | tee gpurun_out/bench_ours_2_pp.log
echo "== bench reference N=2"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29605 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | grep -v "Training Epoch" | tail -3 | tee gpurun_out/bench_ref_2.log
