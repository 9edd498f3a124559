#!/bin/bash
# ncu launch lists (gpu__time_duration.sum) of the one-config bench commands: which kernels a step of C4 / C5 / C3 consists of
TAG=${1:-ll}
mkdir -p gpurun_out
for c in c4 c5 c3; do
  case $c in
    c4) A="--config c4 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline";;
    c5) A="--config c5 --seconds 2 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline";;
    c3) A="--config c3 --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline";;
  esac
  timeout 300 python bench.py $A > gpurun_out/plain_ll_${c}_$TAG.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${c}_$TAG.csv python bench.py $A > gpurun_out/ncu_ll_${c}_$TAG.log 2>&1
  echo "$c exit $?"
done
