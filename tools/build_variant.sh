#!/bin/bash
# tools/build_variant.sh <name> <decode_persistent source>: a copy of the library with another decode_persistent.cu, for A/B runs in ONE
# gpurun call; VARIANT_FLAGS=-DBEVGEN_DP_TRACE builds the clock-trace probes in (tools/decode_trace.py); (boxes differ by ~8 % in sustained clock): BEVGEN_B200_LIB=bevgen_b200/variants/lib_<name>.so python tools/decode_debug.py ...
set -e
cd "$(dirname "$0")/.."
mkdir -p bevgen_b200/variants bevgen_b200/build
cp "$2" bevgen_b200/csrc/_variant_tmp.cu
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $VARIANT_FLAGS -c bevgen_b200/csrc/_variant_tmp.cu -o bevgen_b200/build/_variant_$1.o
rm bevgen_b200/csrc/_variant_tmp.cu
OBJS=$(ls bevgen_b200/build/*.cu.o | grep -v decode_persistent.cu.o)
nvcc -shared -o bevgen_b200/variants/lib_$1.so $OBJS bevgen_b200/build/_variant_$1.o -Xcompiler -fPIC -cudart static
echo built bevgen_b200/variants/lib_$1.so
