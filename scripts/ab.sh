#!/bin/bash
# usage: scripts/ab.sh "<variant> <block> <delta> [ctas_per_sm]" ...   (variant = name under lib/variants or "main")
for cfg in "$@"; do
  set -- $cfg
  if [ "$1" = "main" ]; then unset DTL_LIB; else export DTL_LIB=$PWD/dlsim-mrm_b200/lib/variants/$1.so; fi
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --block $2 --delta $3 --ctas ${4:-0} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1 block $2 delta $3 ctas ${4:-0}:', round(d['value'],2), 'iter/s gteps', round(d['sssp_gteps'],2), 'e2e', round(d['e2e']['value'],2), {k:round(v,2) for k,v in d['phase_ms_per_step'].items() if v>0}, 'relax', int(d['work_per_step']['relaxations']/1e6), 'rounds', int(d['work_per_step']['rounds']/2000), 'reruns', d['work_per_step'].get('reruns'))"
done
