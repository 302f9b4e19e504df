#!/usr/bin/env bash
# A/B runs of the bench step on one GPU: each argument is "label|ENV=val ENV=val|extra bench args"
# prints iterations/s and the per-phase milliseconds of every configuration (4 timed iterations from iteration 3)
mkdir -p gpurun_out
for cfg in "$@"; do
  IFS='|' read -r label envs extra <<< "$cfg"
  env $envs timeout 300 python bench.py --steps ${AB_STEPS:-4} --warmup ${AB_WARMUP:-3} --no-cpu-baseline $extra 2> gpurun_out/ab_last.err | python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
    w=d['work_per_step']
    print('$label:', round(d['value'],2), 'iter/s | e2e', round(d['e2e']['value'],1), '|', {k:round(v,3) for k,v in d['phase_ms_per_step'].items() if v>0}, '| pops', round(w['pops']/1e6,2), 'M windows', int(w['rounds']), 'reruns', w.get('reruns'), 'gap', d['relative_gap_last'], d['roofline']['kernel'][:40])
except Exception as ex:
    print('$label: FAILED', ex)
" 2>&1 | tee -a gpurun_out/ab.log
done
