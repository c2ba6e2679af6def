#!/bin/bash
# Run under gpurun: ncu launch list of one steady-state VQGAN step, one `--set full` capture of the dominant kernel (conv_fused, the
# 128->128 3x3 convs at 256x256), and the same for the stage-2 forward (fused attention).  Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
L=${LAUNCHES_PER_STEP:-223}
BENCH_LITE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L+1)) -c $L --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
F=$(grep -c conv_fused gpurun_out/launches.csv)
echo "conv_fused launches per step: $F"
BENCH_LITE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fused -s $((3*F)) -c 3 \
   -f -o gpurun_out/prof_conv_fused python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fused -s 30 -c 2 \
   -f -o gpurun_out/prof_attn_fused python tools/stage2_perf.py fp32x3 > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out
