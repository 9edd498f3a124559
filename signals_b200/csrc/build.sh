#!/bin/sh
# Builds signals_b200/libsigb200.so for sm_100a (B200).  No torch, no CMake: nvcc only.
# Incremental: a source is recompiled when it, any header here or include/sigb200.h, or this script is newer than
# its object (SIGB_REBUILD=1 forces everything).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$ROOT/signals_b200/libsigb200.so"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I$ROOT/include -I$HERE ${SIGB_NVCC_EXTRA}"
mkdir -p "$HERE/_obj"
STAMP="$HERE/_obj/.flags"
if [ "$(cat "$STAMP" 2>/dev/null)" != "$FLAGS" ]; then SIGB_REBUILD=1; fi
stale() {   # $1 = source, $2 = object
    [ -n "$SIGB_REBUILD" ] && return 0
    [ -f "$2" ] || return 0
    for dep in "$1" "$HERE"/*.h "$HERE"/*.cuh "$ROOT/include/sigb200.h" "$HERE/build.sh"; do
        [ "$dep" -nt "$2" ] && return 0
    done
    return 1
}
PIDS=""
OBJS=""
for src in sigb_kernels.cu sigb_fused.cu sigb_pipe.cu sigb_reg.cu sigb_plan.cu sigb_design.cpp; do
    obj="$HERE/_obj/${src%.*}.o"
    OBJS="$OBJS $obj"
    if stale "$HERE/$src" "$obj"; then
        "$NVCC" $FLAGS -x cu -c "$HERE/$src" -o "$obj" &
        PIDS="$PIDS $!"
    fi
done
for pid in $PIDS; do
    wait "$pid" || { echo "build failed" >&2; exit 1; }
done
printf '%s' "$FLAGS" > "$STAMP"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $OBJS -cudart static
echo "built $OUT"
