#!/usr/bin/env bash
# usage: ab_k1.sh <variant>...   -- K1 (and the iteration) per variant of the library, config 5 on one GPU; "main" = the in-tree library
for v in "$@"; do
  ( [ $v != main ] && export DTL_LIB=$PWD/dlsim-mrm_b200/lib/variants/$v.so
    for rep in 1 2; do echo "variant=$v $(timeout 300 python scripts/shard_step.py 1 12 2>&1 | tail -1 | cut -c1-220)"; done )
done
