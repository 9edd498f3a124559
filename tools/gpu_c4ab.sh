#!/bin/bash
# C4 A/B on one B200: cascade parity tests, then bench --config c4 with reg_variant 0 (delta form) and 4 (state-variable form), alternating.
TAG=${1:-c4ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -s -k "cascade or odd_order or full_size_cascade or fullsize or full_c4" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_$TAG.log | cut -c1-300
grep -E "cascade reg|segments:|C4 full" gpurun_out/pytest_$TAG.log | cut -c1-200
for rep in 1 2; do
for V in 0 4; do
  timeout 600 python bench.py --config c4 --steps 2 --e2e-steps 0 --no-cpu-baseline --plan-opt reg_variant=$V > gpurun_out/bench_c4_${TAG}_v${V}_$rep.json 2> gpurun_out/bench_c4_${TAG}_v${V}_$rep.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c4_${TAG}_v${V}_$rep.json').read().strip().splitlines()[-1])
print('variant $V rep $rep', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('fma_lane_ops_per_clk_sm'), d['clocks'])
PY
done
done
tools/fma_probe > gpurun_out/fma_probe_$TAG.txt 2>&1; grep -E "DELTA|5-coef, R=8|DF2T" gpurun_out/fma_probe_$TAG.txt
