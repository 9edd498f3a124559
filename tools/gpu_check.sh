#!/bin/bash
# One gpurun call on one B200: GPU parity suite, parity table, smoke, then the default bench line.
#   gpurun --timeout 1800 -- 'bash tools/gpu_check.sh [tag]'
# Companions: tools/gpu_record.sh (bench + ncu launch list + full capture of the dominant kernel),
#             tools/gpu_multi.sh N (N ranks: the default bench line with the NCCL reduce and the N-vs-1 parity),
#             tools/gpu_prof.sh tag "c3:k_bank c5:k_voices ..." (ncu --set full of one kernel per config).
TAG=${1:-check}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log | cut -c1-300
timeout 600 python tests/parity_report.py > gpurun_out/parity_$TAG.txt 2> gpurun_out/parity_$TAG.err; echo "parity exit $?"
tail -6 gpurun_out/parity_$TAG.txt
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | tail -3; echo "bench exit $?"
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('C2', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
for k,v in d.get('extra',{}).items():
    if k=='c1':
        for r in v['rows']:
            if r['mode']=='graph': print('c1', r['graph'][:20], r['frames'], round(r['p50_us'],1), round(r['p99_us'],1))
    else:
        print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], v.get('reduce_ms'), v.get('parity_n_vs_1',{}).get('max_abs'), v.get('vs_unmodulated_c2'))
PY
