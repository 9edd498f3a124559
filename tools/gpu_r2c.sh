#!/bin/bash
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$TAG.log
tail -6 gpurun_out/pytest_$TAG.log
grep -h "segments: warm_rows" gpurun_out/pytest_$TAG.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "time_segments_match_oracle" 2>&1 | grep "segments:"
for V in 0 2 3; do
  timeout 300 python bench.py --config c4 --steps 2 --e2e-steps 0 --no-cpu-baseline --plan-opt reg_variant=$V > gpurun_out/c4_v$V.json 2>gpurun_out/c4_v$V.err
  python -c "import json;d=json.loads(open('gpurun_out/c4_v$V.json').read());print('c4 reg_variant $V', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('fma_lane_ops_per_clk_sm'), d['clocks'])"
done
timeout 300 python bench.py --config c1 --no-cpu-baseline > gpurun_out/c1_$TAG.json 2>gpurun_out/c1_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c1_$TAG.json').read().strip().splitlines()[-1])
for r in d['c1']['rows']: print('c1', r['graph'][:20], r['frames'], r['mode'], round(r['p50_us'],1), round(r['p99_us'],1), r['cuda_graph_launches'])
PY
timeout 300 python bench.py --config c5 --voices 131072 --steps 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/c5_131k_$TAG.json 2>/dev/null
python -c "import json;d=json.loads(open('gpurun_out/c5_131k_$TAG.json').read());print('c5 131k default', d['value'], d['ms_per_step'])"
