#!/bin/bash
# phase cycles of the origin-bundle kernel (thread 0 of every CTA) and the per-kernel durations of one 2000-tree launch
mkdir -p gpurun_out
export DTL_SSSP_KERNEL=bundle DTL_BUNDLE_B=${1:-16} DTL_BUNDLE_BLOCK=${2:-1024}
DTL_LIB=$PWD/dlsim-mrm_b200/lib/variants/prof.so DTL_SSSP_PROFILE=1 timeout 300 python scripts/profile_sssp.py 2000 > gpurun_out/bundle_prof.log 2>&1
cat gpurun_out/bundle_prof.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bundle_launches.csv python scripts/profile_sssp.py 2000 > gpurun_out/bundle_ncu.log 2>&1
grep -E "bundle|finish|count" gpurun_out/bundle_launches.csv | cut -d, -f5,12- | head -20
