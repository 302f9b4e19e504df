#!/bin/bash
# K1 launch time for the tree counts of an origin shard (1000 / 500 / 250 trees): default kernels against the origin-bundle kernel
mkdir -p gpurun_out; : > gpurun_out/bundle_shard.log
for n in 1000 500 250; do
  for cfg in "auto 0 0" "bundle 16 1024" "bundle 8 1024" "bundle 8 512" "bundle 4 1024" "bundle 4 512"; do
    set -- $cfg
    if [ "$1" = "auto" ]; then unset DTL_SSSP_KERNEL; else export DTL_SSSP_KERNEL=$1 DTL_BUNDLE_B=$2 DTL_BUNDLE_BLOCK=$3 DTL_BUNDLE_LPT=2; fi
    echo "n=$n $cfg: $(timeout 200 python scripts/profile_sssp.py $n 2>&1 | tail -1 | cut -c1-200)" | tee -a gpurun_out/bundle_shard.log
  done
done
