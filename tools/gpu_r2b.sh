#!/bin/bash
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$TAG.log
tail -6 gpurun_out/pytest_$TAG.log
timeout 600 python tests/parity_report.py > gpurun_out/parity_$TAG.txt 2> gpurun_out/parity_$TAG.err; echo "parity exit $?"
cat gpurun_out/parity_$TAG.txt
for P in 4 8 16; do
  timeout 300 python bench.py --config c5 --voices 131072 --steps 5 --e2e-steps 0 --no-cpu-baseline --plan-opt voices_pieces=$P > gpurun_out/c5_131k_p$P.json 2>gpurun_out/c5_131k_p$P.err
  python -c "import json;d=json.loads(open('gpurun_out/c5_131k_p$P.json').read());print('c5 131k pieces $P', d['value'], d['ms_per_step'])"
  timeout 300 python bench.py --config c5 --steps 3 --e2e-steps 0 --no-cpu-baseline --plan-opt voices_pieces=$P > gpurun_out/c5_1m_p$P.json 2>/dev/null
  python -c "import json;d=json.loads(open('gpurun_out/c5_1m_p$P.json').read());print('c5 1M pieces $P', d['value'], d['ms_per_step'])"
done
timeout 300 python bench.py --config c1 --no-cpu-baseline > gpurun_out/c1_$TAG.json 2>gpurun_out/c1_$TAG.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c1_r02b.json').read().strip().splitlines()[-1])
for r in d['c1']['rows']: print('c1', r['graph'][:20], r['frames'], r['mode'], round(r['p50_us'],1), round(r['p99_us'],1), r['cuda_graph_launches'])
PY
