#!/bin/bash
# Round-2 record: default bench (headline C2 + extra), launch list of the same command, full capture of the dominant kernel.
TAG=${1:-record}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | tail -3; echo "bench exit $?"
kill $SMI
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('C2', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e'].get('render_host_gbs_this_rank'), d['e2e'].get('plain_d2h_memcpy_gbs_this_rank_all_ranks_copying'))
for k,v in d.get('extra',{}).items():
    if k!='c1': print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], v.get('vs_unmodulated_c2'))
PY
BCMD="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-extra"
timeout 300 $BCMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $BCMD > gpurun_out/ncu_l_$TAG.log 2>&1
timeout 300 $BCMD > gpurun_out/plain2_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain_scan3 -s 3 -c 1 -f -o gpurun_out/prof_scan3_$TAG $BCMD > gpurun_out/ncu_f_$TAG.log 2>&1
echo "ncu exit $?"
