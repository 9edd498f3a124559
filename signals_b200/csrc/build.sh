#!/bin/sh
# Builds signals_b200/libsigb200.so for sm_100a (B200).  No torch, no CMake: nvcc only.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$ROOT/signals_b200/libsigb200.so"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I$ROOT/include -I$HERE ${SIGB_NVCC_EXTRA}"
mkdir -p "$HERE/_obj"
PIDS=""
for src in sigb_kernels.cu sigb_fused.cu sigb_pipe.cu sigb_reg.cu sigb_plan.cu; do
    "$NVCC" $FLAGS -c "$HERE/$src" -o "$HERE/_obj/${src%.cu}.o" &
    PIDS="$PIDS $!"
done
"$NVCC" $FLAGS -x cu -c "$HERE/sigb_design.cpp" -o "$HERE/_obj/sigb_design.o" &
PIDS="$PIDS $!"
for pid in $PIDS; do
    wait "$pid" || { echo "build failed" >&2; exit 1; }
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$HERE/_obj/sigb_kernels.o" "$HERE/_obj/sigb_fused.o" \
    "$HERE/_obj/sigb_pipe.o" "$HERE/_obj/sigb_reg.o" "$HERE/_obj/sigb_plan.o" "$HERE/_obj/sigb_design.o" -cudart static
echo "built $OUT"
