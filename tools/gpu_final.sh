#!/bin/bash
# Round-end check: full GPU parity suite, smoke, bench lines (C2 default, C4), ncu launch list + full capture of k_cascade_reg.
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke_$TAG.log
tail -2 gpurun_out/smoke_$TAG.log
for c in c2 c4; do
  timeout 600 python bench.py --config $c > gpurun_out/bench_${c}_$TAG.json 2> gpurun_out/bench_${c}_$TAG.err; echo "bench $c exit $?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${c}_$TAG.json').read().strip().splitlines()[-1])
    print('$c', 'value %.4g'%d['value'], 'roofline', d['roofline']['bound'], '%.3f'%d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'ms/step %.3f'%d['ms_per_step'], d['clocks'])
except Exception as e: print('$c parse failed', e)
PY
  tail -2 gpurun_out/bench_${c}_$TAG.err
done
BCMD="python bench.py --config c4 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c4_$TAG.csv $BCMD > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cascade_reg -s 20 -c 1 -f -o gpurun_out/prof_reg_$TAG $BCMD > gpurun_out/ncu_f_$TAG.log 2>&1; echo "ncu full exit $?"
for c in c5 c3; do
  timeout 600 python bench.py --config $c > gpurun_out/bench_${c}_$TAG.json 2> gpurun_out/bench_${c}_$TAG.err; echo "bench $c exit $?"
  cut -c1-200 gpurun_out/bench_${c}_$TAG.json
done
