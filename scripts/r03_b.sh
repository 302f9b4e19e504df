#!/usr/bin/env bash
mkdir -p gpurun_out
t=r03b
for n in 250 148; do
DTL_SSSP_PROFILE=1 DTL_LIB=$PWD/dlsim-mrm_b200/lib/variants/prof.so PROFILE_CONTIGUOUS=1 timeout 300 python scripts/profile_sssp.py $n > gpurun_out/${t}_smq_prof_$n.log 2>&1
tail -3 gpurun_out/${t}_smq_prof_$n.log
done
PROFILE_CONTIGUOUS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:sssp_smq -s 1 -c 1 -f -o gpurun_out/${t}_ncu_smq python scripts/profile_sssp.py 250 > gpurun_out/${t}_ncu_smq.log 2>&1
ncu -i gpurun_out/${t}_ncu_smq.ncu-rep --page raw --csv > gpurun_out/${t}_ncu_smq_raw.csv 2>/dev/null
ncu -i gpurun_out/${t}_ncu_smq.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${t}_ncu_smq_src.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${t}_ncu_smq_raw.csv gpurun_out/${t}_ncu_smq_src.csv > gpurun_out/${t}_ncu_smq.md 2>&1
tail -30 gpurun_out/${t}_ncu_smq.md
