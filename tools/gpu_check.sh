#!/bin/bash
# One gpurun call: GPU parity tests, smoke, default bench, then the ncu launch list and one
# full capture of the dominant kernel (each ncu pass only after the same command exited 0).
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_$TAG.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
cat gpurun_out/bench_$TAG.json
kill $SMI
BCMD="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 300 $BCMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $BCMD > gpurun_out/ncu_l_$TAG.log 2>&1
timeout 300 $BCMD > gpurun_out/plain2_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain_scan3 -s 3 -c 1 -f -o gpurun_out/prof_$TAG $BCMD > gpurun_out/ncu_f_$TAG.log 2>&1
# C4: launch list and full capture of the register-resident cascade kernel (one 10 s slab per launch)
C4CMD="python bench.py --config c4 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 300 $C4CMD > gpurun_out/plain_c4_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cascade_reg -s 20 -c 1 -f -o gpurun_out/prof_reg_$TAG $C4CMD > gpurun_out/ncu_reg_$TAG.log 2>&1
ls -la gpurun_out
