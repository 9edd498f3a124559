#!/bin/bash
# C4 with k_cascade_delta geometries (delta_probe 0..N), alternating, one B200
TAG=${1:-c4p}; shift
mkdir -p gpurun_out
for rep in 1 2; do
for V in "$@"; do
  timeout 600 python bench.py --config c4 --steps 2 --e2e-steps 0 --no-cpu-baseline --plan-opt delta_probe=$V > gpurun_out/bench_c4_${TAG}_p${V}_$rep.json 2> gpurun_out/bench_c4_${TAG}_p${V}_$rep.err || tail -3 gpurun_out/bench_c4_${TAG}_p${V}_$rep.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c4_${TAG}_p${V}_$rep.json').read().strip().splitlines()[-1])
print('delta_probe $V rep $rep', '%.4g' % d['value'], round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done
done
