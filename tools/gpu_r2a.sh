#!/bin/bash
# Round 2, call A: GPU tests, smoke, default bench (with extra), k_voices at one GPU's share of C5 on 8 GPUs.
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$TAG.log
tail -15 gpurun_out/pytest_$TAG.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke_$TAG.log
tail -2 gpurun_out/smoke_$TAG.log
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | tail -3; echo "bench exit $?"
tail -5 gpurun_out/bench_$TAG.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02a.json').read().strip().splitlines()[-1])
print('C2', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e'].get('render_host_gbs_this_rank'))
for k,v in d.get('extra',{}).items():
    if k=='c1':
        for r in v['rows']: print('c1', r['graph'][:20], r['frames'], r['mode'], r['p50_us'], r['p99_us'], r['cuda_graph_launches'])
        print(v.get('cpu_reference_blockwise'))
    else:
        print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], v.get('reduce_ms'), v.get('parity_n_vs_1'))
PY
for P in 1 2 4; do
  timeout 300 python bench.py --config c5 --voices 131072 --steps 5 --e2e-steps 0 --no-cpu-baseline --plan-opt voices_pieces=$P > gpurun_out/c5_131k_p$P.json 2>gpurun_out/c5_131k_p$P.err
  python -c "import json;d=json.loads(open('gpurun_out/c5_131k_p$P.json').read());print('c5 131k pieces $P', d['value'], d['ms_per_step'])"
done
timeout 300 python bench.py --config c5 --steps 3 --e2e-steps 0 --no-cpu-baseline --plan-opt voices_pieces=1 > gpurun_out/c5_1m_p1.json 2>/dev/null
python -c "import json;d=json.loads(open('gpurun_out/c5_1m_p1.json').read());print('c5 1M pieces 1', d['value'], d['ms_per_step'])"
timeout 300 python bench.py --config c5 --steps 3 --e2e-steps 0 --no-cpu-baseline --plan-opt voices_pieces=4 > gpurun_out/c5_1m_p4.json 2>/dev/null
python -c "import json;d=json.loads(open('gpurun_out/c5_1m_p4.json').read());print('c5 1M pieces 4', d['value'], d['ms_per_step'])"
