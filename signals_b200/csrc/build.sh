#!/bin/sh
# Builds signals_b200/libsigb200.so for sm_100a (B200).  No torch, no CMake: nvcc only.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$ROOT/signals_b200/libsigb200.so"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I$ROOT/include -I$HERE"
mkdir -p "$HERE/_obj"
"$NVCC" $FLAGS ${SIGB_NVCC_EXTRA} -c "$HERE/sigb_kernels.cu" -o "$HERE/_obj/sigb_kernels.o" &
"$NVCC" $FLAGS ${SIGB_NVCC_EXTRA} -c "$HERE/sigb_fused.cu" -o "$HERE/_obj/sigb_fused.o" &
"$NVCC" $FLAGS ${SIGB_NVCC_EXTRA} -c "$HERE/sigb_pipe.cu" -o "$HERE/_obj/sigb_pipe.o" &
"$NVCC" $FLAGS -c "$HERE/sigb_plan.cu" -o "$HERE/_obj/sigb_plan.o" &
"$NVCC" $FLAGS -x cu -c "$HERE/sigb_design.cpp" -o "$HERE/_obj/sigb_design.o" &
wait
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$HERE/_obj/sigb_kernels.o" "$HERE/_obj/sigb_fused.o" "$HERE/_obj/sigb_pipe.o" "$HERE/_obj/sigb_plan.o" "$HERE/_obj/sigb_design.o" -cudart static
echo "built $OUT"
