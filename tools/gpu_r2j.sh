#!/bin/bash
TAG=${1:-r02j}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log | cut -c1-300
for V in 131072 1048576; do
  timeout 300 python bench.py --config c5 --voices $V --steps 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/c5_${V}_$TAG.json 2>/dev/null
  python -c "import json;d=json.loads(open('gpurun_out/c5_${V}_$TAG.json').read());print('c5 $V', d['value'], d['ms_per_step'])"
done
