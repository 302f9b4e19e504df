#!/usr/bin/env bash
mkdir -p gpurun_out
t=r03c
log=gpurun_out/${t}_smq_variants.log; : > $log
DTL_SSSP_KERNEL=smq timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu 2>&1 | tail -2 | tee -a $log
for world in 8 4; do
  for v in main smq_old smq_b smq_g1u4 smq_g4u1; do
    ( [ $v != main ] && export DTL_LIB=$PWD/dlsim-mrm_b200/lib/variants/$v.so
      echo "variant=$v $(timeout 200 python scripts/shard_step.py $world 9 2>&1 | tail -1 | cut -c1-260)" ) | tee -a $log
  done
done
for d in 0.5 2 4; do
  ( export DTL_SSSP_SMQ_DELTA=$d; echo "delta x$d $(timeout 200 python scripts/shard_step.py 8 9 2>&1 | tail -1 | cut -c1-260)" ) | tee -a $log
done
for b in 256 1024; do
  ( export DTL_SSSP_SMQ_BLOCK=$b; echo "block $b $(timeout 200 python scripts/shard_step.py 8 9 2>&1 | tail -1 | cut -c1-260)" ) | tee -a $log
done
for b in 512 1024; do
  ( export DTL_SSSP_SMQ_BLOCK=$b; echo "world 4 block $b $(timeout 200 python scripts/shard_step.py 4 9 2>&1 | tail -1 | cut -c1-260)" ) | tee -a $log
done
