#!/bin/bash
TAG=${1:-r02d}
mkdir -p gpurun_out
for V in 0 3; do for P in 1 2 3; do
  timeout 300 python bench.py --config c4 --steps 2 --e2e-steps 0 --no-cpu-baseline --plan-opt reg_variant=$V --plan-opt reg_pieces=$P > gpurun_out/c4_v${V}_p$P.json 2>gpurun_out/c4_v${V}_p$P.err
  python -c "import json;d=json.loads(open('gpurun_out/c4_v${V}_p$P.json').read());print('c4 reg_variant $V pieces $P', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('fma_lane_ops_per_clk_sm'), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
