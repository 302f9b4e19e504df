#!/bin/bash
# usage: scripts/ncu_sssp.sh <tag> <n_trees> <block> <delta> [ctas]   -- ncu --set full of one SSSP launch (second launch of profile_sssp.py)
tag=$1; shift
ncu --set full --import-source on --clock-control none -k regex:sssp_kernel -s 1 -c 1 -f -o gpurun_out/ncu_$tag python scripts/profile_sssp.py "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$tag.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/ncu_${tag}_src.csv 2>/dev/null
tail -2 gpurun_out/ncu_$tag.log
