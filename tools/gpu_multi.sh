#!/bin/bash
# N ranks on N GPUs: the default bench line (C2 weak + extra.c5 strong with the NCCL reduce and the N-vs-1 parity)
N=${1:-8}; TAG=${2:-multi}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/bench${N}_$TAG.json 2> gpurun_out/bench${N}_$TAG.err ) 2>&1 | tail -3
echo "bench exit $?"; tail -3 gpurun_out/bench${N}_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench${N}_$TAG.json').read().strip().splitlines()[-1])
print('C2', d['n_gpus'], d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e'].get('render_host_gbs_this_rank'), d['e2e'].get('pinned_numa_node'), d['e2e'].get('cpus_bound'))
c5=d['extra']['c5']
print('c5', c5['value'], c5['ms_per_step'], c5.get('reduce_ms'), c5.get('parity_n_vs_1'))
PY
