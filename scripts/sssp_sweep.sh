#!/bin/bash
# usage: scripts/sssp_sweep.sh "<variant> <n_trees> <block> <smem_cap> <delta>" ...  -- one SSSP launch per config on the config-5 grid
for cfg in "$@"; do
  set -- $cfg
  if [ "$1" = "main" ]; then unset DTL_LIB; else export DTL_LIB=$PWD/dlsim-mrm_b200/lib/variants/$1.so; fi
  DTL_SSSP_PROFILE=1 DTL_SSSP_ICAP=$4 python scripts/profile_sssp.py $2 $3 $5 2>&1 | grep -v "^0 " | sed "s/^/$1 n=$2 block=$3 cap=$4 delta=$5: /"
done
