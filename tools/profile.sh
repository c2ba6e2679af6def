#!/bin/bash
# Run under gpurun: ncu launch list of one steady-state VQGAN step (only this library's kernels: -k regex over their names; the last
# quarter of the 4 captured steps is the steady-state one), plus `--set full` captures of the dominant kernel.  Numbers printed under ncu
# are never bench values.
set -x
mkdir -p gpurun_out
P=${PRECISION:-f16f8}
K='regex:^(attn_fused|attn_softmax|conv_fused2|conv_fused3|conv_fused|conv_halo|conv_in3|conv_out3|to_uint8_hwc|denorm|gather_rows|gemm_tc|gn_affine|gn_finalize|gn_stats|im2col3x3|prep|row_sqnorm|softmax_rows|transpose|vq_nearest)_kernel'
BENCH_LITE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 2000 --csv \
   --log-file gpurun_out/launches_all.csv python bench.py --precision $P --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
lines = [l for l in open("gpurun_out/launches_all.csv") if not l.startswith("==")]
hdr, rows = lines[0], lines[1:]
n = len(rows) // 4
open("gpurun_out/launches_step.csv", "w").writelines([hdr] + rows[-n:])
print("kernels per step:", n)
PY
python tools/summarize_launches.py gpurun_out/launches_step.csv
# `--set full` of three consecutive dominant launches (conv 128->128 3x3 at 256x256 x 96: conv1 without / conv2 with residual) inside a step
# (the conv_fused2 full capture is optional: set FULL=1)
F=$(grep -c conv_fused2 gpurun_out/launches_step.csv)
if [ "${FULL:-0}" = "1" ]; then
BENCH_LITE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fused2 -s $((3*F+2)) -c 3 \
   -f -o gpurun_out/prof_conv_fused python bench.py --precision $P --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
fi
