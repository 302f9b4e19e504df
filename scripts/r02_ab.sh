#!/usr/bin/env bash
# A/B runs of the bench step on one GPU: each argument is "label|ENV=val ENV=val|extra bench args"
# prints iterations/s and the per-phase milliseconds of every configuration
mkdir -p gpurun_out
for cfg in "$@"; do
  IFS='|' read -r label envs extra <<< "$cfg"
  env $envs timeout 300 python bench.py --steps ${AB_STEPS:-20} --warmup ${AB_WARMUP:-5} --no-cpu-baseline $extra 2> gpurun_out/ab_last.err | python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
    w=d['work_per_step']
    print('$label:', round(d['value'],2), 'iter/s | e2e', round(d['e2e']['value'],1), '|', {k:round(v,3) for k,v in d['phase_ms_per_step'].items() if v>0}, '| gap', d['relative_gap_last'], 'digest ok', d['volume_digest']['matches_single_gpu_golden'])
except Exception as ex:
    print('$label: FAILED', ex)
" 2>&1 | tee -a gpurun_out/ab.log
done
