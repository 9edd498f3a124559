#!/bin/bash
TAG=${1:-r02i}
mkdir -p gpurun_out
CMD="python bench.py --config c5 --voices 131072 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_c5_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_voices -s 6 -c 1 -f -o gpurun_out/prof_voices131k_$TAG $CMD > gpurun_out/ncu_voices_$TAG.log 2>&1
echo "ncu exit $?"
tail -2 gpurun_out/plain_c5_$TAG.log | cut -c1-300
