#!/bin/bash
# one gpurun call: parity tests of the K1 kernels, then A/B of the origin-bundle kernel against the global-queue kernel
# usage: scripts/bundle_ab.sh "<kernel> <B> <block> <delta factor> <origins per thread> [library variant under lib/variants]" ...
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q > gpurun_out/pytest_engine.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_engine.log
tail -5 gpurun_out/pytest_engine.log
: > gpurun_out/bundle_ab.log
for cfg in "$@"; do
  set -- $cfg
  export DTL_SSSP_KERNEL=$1 DTL_BUNDLE_B=$2 DTL_BUNDLE_BLOCK=$3 DTL_BUNDLE_LPT=${5:-4}
  if [ -n "$6" ]; then export DTL_LIB=$PWD/dlsim-mrm_b200/lib/variants/$6.so; else unset DTL_LIB; fi
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --delta $4 2>gpurun_out/bundle_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$cfg:', round(d['value'],2), 'iter/s gteps', round(d['sssp_gteps'],2), 'e2e', round(d['e2e']['value'],2), {k:round(v,2) for k,v in d['phase_ms_per_step'].items() if v>0}, 'relax', int(d['work_per_step']['relaxations']/1e6), 'pops', int(d['work_per_step']['pops']/1e6), 'rounds', int(d['work_per_step']['rounds']), 'reruns', d['work_per_step'].get('reruns'), 'gap', d['relative_gap_last'])" 2>&1 | tee -a gpurun_out/bundle_ab.log
done
