#!/bin/bash
# Iteration run: GPU tests, then one bench line per config.  Usage: gpurun -- 'bash tools/gpu_iter.sh tag [configs]'
TAG=${1:-it}
CONFIGS=${2:-"c2 c3 c4 c5"}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$TAG.log
tail -15 gpurun_out/pytest_$TAG.log
for c in $CONFIGS; do
  timeout 600 python bench.py --config $c > gpurun_out/bench_${c}_$TAG.json 2> gpurun_out/bench_${c}_$TAG.err; echo "bench $c exit $?"
  cut -c1-250 gpurun_out/bench_${c}_$TAG.json; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${c}_$TAG.json').read().strip().splitlines()[-1])
    print('$c', 'value %.4g'%d['value'], 'roofline', d['roofline']['bound'], '%.3f'%d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'ms/step %.3f'%d['ms_per_step'])
except Exception as e: print('$c parse failed', e)
PY
  tail -3 gpurun_out/bench_${c}_$TAG.err
done
