#!/bin/bash
# C4 iteration: cascade tests, then bench lines of the register-resident cascade kernel against the pipelined one.
TAG=${1:-c4}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "cascade or segments or full_size_cascade" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$TAG.log
grep -E "cascade reg|pipe segments|passed|failed|Error|assert" gpurun_out/pytest_$TAG.log | tail -30
B="timeout 300 python bench.py --config c4 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
run() { name=$1; shift; $B "$@" > gpurun_out/bench_c4_${TAG}_$name.json 2> gpurun_out/bench_c4_${TAG}_$name.err; echo "$name exit $?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c4_${TAG}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'ms/step %.3f'%d['ms_per_step'], d.get('clocks'))
except Exception as e: print('$name parse failed', e)
PY
}
run reg0
run reg1 --plan-opt reg_variant=1
run reg0_10s --slab-seconds 10
run reg1_10s --slab-seconds 10 --plan-opt reg_variant=1
