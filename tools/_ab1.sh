mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_persistent_gpu.py tests/test_decode_gpu.py -x -q 2>&1 | tail -2
for v in new prev new prev new prev new prev; do
  if [ $v = new ]; then unset BEVGEN_B200_LIB; else export BEVGEN_B200_LIB=bevgen_b200/variants/lib_$v.so; fi
  DP_PROFILE=0 timeout 180 python tools/decode_debug.py plain 1536 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v', d.get('run0_s'), d.get('run1_s'), d.get('ok'), d.get('failure(code,cta,step,layer,phase,a,b,thread)'))" | tee -a gpurun_out/tag_ab10.log
done
