#!/bin/bash
# shard-sized K1 launches: asynchronous origin bundles of 2 / 4 / 8 origins against the default kernel choice (spread origins)
mkdir -p gpurun_out; : > gpurun_out/bundle_shard3.log
export DTL_BUNDLE_ASYNC=1
for n in 250 500 1000; do
  for cfg in "auto 0" "bundle 2" "bundle 4" "bundle 8"; do
    set -- $cfg
    if [ "$1" = "auto" ]; then unset DTL_SSSP_KERNEL DTL_BUNDLE_B; else export DTL_SSSP_KERNEL=$1 DTL_BUNDLE_B=$2; fi
    echo "n=$n $cfg: $(timeout 200 python scripts/profile_sssp.py $n 2>&1 | tail -1 | cut -c1-200)" | tee -a gpurun_out/bundle_shard3.log
  done
done
