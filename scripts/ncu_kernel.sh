#!/bin/bash
# usage: scripts/ncu_kernel.sh <tag> <kernel regex> <python script and args...>   -- ncu --set full of the second matching launch
tag=$1; rx=$2; shift 2
ncu --set full --import-source on --clock-control none -k regex:$rx -s 1 -c 1 -f -o gpurun_out/ncu_$tag python "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$tag.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/ncu_${tag}_src.csv 2>/dev/null
tail -2 gpurun_out/ncu_$tag.log
