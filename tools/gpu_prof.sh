#!/bin/bash
# ncu --set full of one kernel per config.  Usage: gpurun -- 'bash tools/gpu_prof.sh tag "c4:k_chain_scan2 c3:k_bank c5:k_voices"'
TAG=${1:-p}
shift
mkdir -p gpurun_out
for spec in $1; do
  c=${spec%%:*}; k=${spec##*:}
  case $c in
    c2) A="--config c2 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline $C2_EXTRA";;
    c3) A="--config c3 --seconds 2 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline";;
    c4) A="--config c4 --seconds 10 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline";;
    c5) A="--config c5 --seconds 1 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline";;
  esac
  timeout 300 python bench.py $A > gpurun_out/plain_${c}_$TAG.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${c}_$TAG python bench.py $A > gpurun_out/ncu_${c}_$TAG.log 2>&1
  echo "$c $k exit $?"; tail -2 gpurun_out/ncu_${c}_$TAG.log
done
ls -la gpurun_out | tail -8
