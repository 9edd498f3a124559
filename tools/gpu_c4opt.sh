#!/bin/bash
# C4 A/B over one plan option: bash tools/gpu_c4opt.sh TAG KEY V1 V2 ...   (alternating, two repetitions; cascade parity tests first)
TAG=$1; KEY=$2; shift; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -s -k "cascade or odd_order or fullsize or full_c4" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"
tail -2 gpurun_out/pytest_$TAG.log | cut -c1-300
for rep in 1 2; do
for V in "$@"; do
  timeout 600 python bench.py --config c4 --steps 2 --e2e-steps 0 --no-cpu-baseline --plan-opt $KEY=$V > gpurun_out/bench_c4_${TAG}_${V}_$rep.json 2> gpurun_out/bench_c4_${TAG}_${V}_$rep.err || tail -3 gpurun_out/bench_c4_${TAG}_${V}_$rep.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c4_${TAG}_${V}_$rep.json').read().strip().splitlines()[-1])
print('$KEY $V rep $rep', '%.4g' % d['value'], round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done
done
