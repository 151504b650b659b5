#!/bin/bash
mkdir -p gpurun_out
echo "== ncu attn bwd"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -c 1 -o gpurun_out/prof_attn_bwd_v3 -f python scripts/ncu_targets.py attn 2>&1 | tail -2
echo "== ncu attn fwd"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -c 1 -o gpurun_out/prof_attn_fwd_v3 -f python scripts/ncu_targets.py attn 2>&1 | tail -2
