#!/bin/bash
# Run under gpurun: bench line, ncu launch list of one steady-state step, one `--set full` capture of the top kernel.
set -x
mkdir -p gpurun_out
python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json
L=$(python -c "import json;d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]);print(d['gpu_launches']//d['steps'])")
G=$(python -c "import json;d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]);print(d['roofline']['launches_per_step'])")
echo "launches/step=$L gemm launches/step=$G"
BENCH_LITE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L+1)) -c $L --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 ${BENCH_ARGS} > gpurun_out/ncu_list.log 2>&1
BENCH_LITE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s $((G+1)) -c 3 \
   -f -o gpurun_out/prof_gemm python bench.py --steps 1 --warmup 3 ${BENCH_ARGS} > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
