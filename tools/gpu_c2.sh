#!/bin/bash
# C2 iteration: k_osc_reg tests, then C2 bench lines with the register-resident oscillator chain against the scan kernel.
TAG=${1:-c2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s -k "osc_reg or oscreg or cascade or segments" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$TAG.log
grep -E "^osc_reg|passed|failed|Error|assert" gpurun_out/pytest_$TAG.log | tail -14
B="timeout 300 python bench.py --config c2 --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
run() { name=$1; shift; $B "$@" > gpurun_out/bench_c2_${TAG}_$name.json 2> gpurun_out/bench_c2_${TAG}_$name.err; echo "$name exit $?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c2_${TAG}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'ms/step %.3f'%d['ms_per_step'], d.get('clocks'))
except Exception as e: print('$name parse failed', e)
PY
}
run scan3
run oscreg --plan-opt osc_reg=1
run oscreg_seg128 --plan-opt osc_reg=1 --plan-opt pipe_segments=128
