#!/bin/bash
# Round-2, 8-GPU box (8x charge: keep it short).  Run AFTER scripts/validate_round2.sh passed on 2 GPUs.
#   gpurun --gpus 8 --timeout 480 -- 'bash scripts/validate_round2_8gpu.sh'
mkdir -p gpurun_out
run() { n=$1; port=$2; shift 2
  timeout --signal=KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" 2>&1 | grep '^{"metric\|Error\|error\|Traceback' | tail -3 | cut -c1-2000; }
echo "== ours N=8 default";                    run 8 29901 --steps 4 --warmup 3 | tee gpurun_out/bench_ours_8_r2.log
echo "== ours N=8 + CTA-pair fused TP";        QN_TP_2CTA=1 run 8 29902 --steps 4 --warmup 3 --no-e2e | tee gpurun_out/bench_ours_8_tp2cta.log
echo "== ours N=8 + interleaved V=2";          run 8 29903 --steps 4 --warmup 3 --no-e2e --virtual-stages 2 | tee gpurun_out/bench_ours_8_vp2.log
echo "== ours N=8 + both";                     QN_TP_2CTA=1 run 8 29904 --steps 4 --warmup 3 --no-e2e --virtual-stages 2 | tee gpurun_out/bench_ours_8_both.log
echo "== ours N=4 default";                    run 4 29905 --steps 4 --warmup 3 --no-e2e | tee gpurun_out/bench_ours_4_r2.log
echo "== ours N=4 + interleaved V=2";          run 4 29906 --steps 4 --warmup 3 --no-e2e --virtual-stages 2 | tee gpurun_out/bench_ours_4_vp2.log
echo "== ViT ours N=8";                        run 8 29907 --steps 10 --warmup 3 --workload vit | tee gpurun_out/bench_vit_ours_8.log
echo "== ViT reference N=8";                   run 8 29908 --steps 10 --warmup 3 --workload vit --impl reference | tee gpurun_out/bench_vit_ref_8.log
