#!/bin/bash
# K1 launch time for shard-sized tree counts: default kernel choice against asynchronous origin bundles, for origins spread over
# the network (zone % world sharding) and for a contiguous block of origins
mkdir -p gpurun_out; : > gpurun_out/bundle_shard2.log
for n in 1000 500 250; do
  for contiguous in "" 1; do
    export PROFILE_CONTIGUOUS=$contiguous
    for cfg in "auto 0" "bundle 16" "bundle 8"; do
      set -- $cfg
      if [ "$1" = "auto" ]; then unset DTL_SSSP_KERNEL DTL_BUNDLE_B; else export DTL_SSSP_KERNEL=$1 DTL_BUNDLE_B=$2; fi
      echo "n=$n contiguous=${contiguous:-0} $cfg: $(timeout 200 python scripts/profile_sssp.py $n 2>&1 | tail -1 | cut -c1-200)" | tee -a gpurun_out/bundle_shard2.log
    done
  done
done
