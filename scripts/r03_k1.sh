#!/usr/bin/env bash
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(b['value'],2), 'K1', round(b['roofline']['kernel_ms_per_launch'],3), 'iter', round(b['ms_per_step'],3), 'pops', round(b['work_per_step']['pops']/1e6,2), 'rounds', round(b['work_per_step']['rounds']))"; }
echo "default:        $(run)"
echo "block 896:      $(DTL_BUNDLE_ASYNC_BLOCK=896 run)"
for d in 9 10.5 12 14 16; do echo "delta $d:   $(run --delta $d)"; done
echo "block 896 d14:  $(DTL_BUNDLE_ASYNC_BLOCK=896 run --delta 14)"
