#!/usr/bin/env bash
# what one rank of a 2/4/8-GPU run does (shard-only mode on one GPU), per K1 kernel choice and sharding rule
mkdir -p gpurun_out; : > gpurun_out/r02_shards.log
for world in 2 4 8; do
  for cfg in "auto" "bundle 16" "bundle 8" "smq" "gq"; do
    set -- $cfg
    for shard in spatial mod; do
      [ "$shard" = "mod" ] && [ "$1" != "auto" ] && [ "$1" != "bundle" ] && continue
      ( if [ "$1" != "auto" ]; then export DTL_SSSP_KERNEL=$1; fi
        if [ -n "$2" ]; then export DTL_BUNDLE_B=$2; fi
        if [ "$shard" = "mod" ]; then export DTL_SHARD=mod; fi
        echo "shard=$shard B=${2:-} $(timeout 200 python scripts/shard_step.py $world 7 2>&1 | tail -1 | cut -c1-220)" ) | tee -a gpurun_out/r02_shards.log
    done
  done
done
