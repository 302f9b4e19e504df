#!/bin/bash
# one gpurun call for the round's evidence: bench (both arms), ncu launch list of the bench command, ncu --set full of every kernel
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
cat gpurun_out/bench_${tag}.json | head -c 600; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_${tag}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_${tag}.log 2>&1
# full capture: one launch of each kernel from a short run (iteration 3 onwards: skip the first 3 iterations' launches)
ncu --set full --import-source on --clock-control none -k regex:'sssp_bundle_async_kernel|sssp_bundle_kernel|bundle_finish_kernel|sssp_kernel|sssp_smq_kernel|backtrace_kernel|scatter_kernel|vdf_kernel|pool_column_cost_kernel|pool_update_kernel' \
    -s 16 -c 10 -f -o gpurun_out/ncu_full_${tag} python scripts/short_loop.py > gpurun_out/ncu_full_${tag}.log 2>&1
ncu -i gpurun_out/ncu_full_${tag}.ncu-rep --page raw --csv > gpurun_out/ncu_full_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_full_${tag}.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/ncu_full_${tag}_src.csv 2>/dev/null
tail -3 gpurun_out/ncu_full_${tag}.log
# DRAM traffic of K1 at the bench configuration (2 000 trees in one launch = bundle search + label->tree pass): roofline.traffic
ncu --set full --import-source on --clock-control none -k regex:'sssp_bundle_async_kernel|sssp_bundle_kernel|bundle_finish_kernel|sssp_kernel' -s 8 -c 2 -f -o gpurun_out/ncu_k1_bench_${tag} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1_bench_${tag}.log 2>&1
ncu -i gpurun_out/ncu_k1_bench_${tag}.ncu-rep --page raw --csv > gpurun_out/ncu_k1_bench_${tag}_raw.csv 2>/dev/null
tail -2 gpurun_out/ncu_k1_bench_${tag}.log
ncu -i gpurun_out/ncu_k1_bench_${tag}.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/ncu_k1_bench_${tag}_src.csv 2>/dev/null
