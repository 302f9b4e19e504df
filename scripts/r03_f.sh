#!/usr/bin/env bash
mkdir -p gpurun_out
t=${1:-r03f}
bash scripts/r03_d.sh $t
DTL_BUNDLE_FILL=1 timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('with the fill:', round(b['value'],2), {k: round(v, 3) for k, v in b['phase_ms_per_step'].items()})"
