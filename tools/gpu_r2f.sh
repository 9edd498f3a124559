#!/bin/bash
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s -k "bank" > gpurun_out/pytest_bank_$TAG.log 2>&1; echo "pytest bank exit $?"
grep -i "max-abs\|passed\|failed\|error" gpurun_out/pytest_bank_$TAG.log | head -20
for U in 1 2; do
  timeout 300 python bench.py --config c3 --steps 5 --e2e-steps 0 --no-cpu-baseline --plan-opt bank_unroll=$U > gpurun_out/c3_u$U.json 2>gpurun_out/c3_u$U.err
  python -c "import json;d=json.loads(open('gpurun_out/c3_u$U.json').read());print('c3 unroll $U', d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
CMD="python bench.py --config c3 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_c3_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bank -s 3 -c 1 -f -o gpurun_out/prof_bank_$TAG $CMD > gpurun_out/ncu_bank_$TAG.log 2>&1
echo "ncu exit $?"
