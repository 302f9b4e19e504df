#!/usr/bin/env bash
# ncu --set full of the four per-iteration kernels at the bench configuration, iteration 9 (scatter, bundle search, label->tree pass, back-trace)
tag=${1:-r02b}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:'scatter_kernel|sssp_bundle_async_kernel|bundle_finish_kernel|backtrace' -s 36 -c 4 -f \
    -o gpurun_out/ncu_${tag} python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${tag}.log 2>&1
tail -3 gpurun_out/ncu_${tag}.log
ncu -i gpurun_out/ncu_${tag}.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_${tag}.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/ncu_${tag}_src.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/ncu_${tag}_raw.csv gpurun_out/ncu_${tag}_src.csv > gpurun_out/ncu_${tag}.md 2>&1
ls -la gpurun_out | tail -8
