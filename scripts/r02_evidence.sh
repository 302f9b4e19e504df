#!/usr/bin/env bash
# round 2 evidence run on one B200: GPU parity tests, the bench line, the launch list of the bench command, and ncu --set full of
# EVERY per-iteration kernel at the bench command (iteration >= 10) plus the column-pool kernels after 20 iterations.
# Each ncu pass runs only after its command has exited 0 without ncu.  tag = $1 (file prefix under gpurun_out/)
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc $?" | tee -a gpurun_out/${tag}_pytest.log
  tail -4 gpurun_out/${tag}_pytest.log
fi
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err; rc=$?; echo "bench rc $rc"
[ $rc -ne 0 ] && exit 1
# launch list (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_ncu_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
# --set full of the per-iteration kernels, 11th occurrence of each (iteration 10)
ncu --set full --import-source on --clock-control none -k regex:'scatter_kernel|sssp_bundle_async_kernel|backtrace|vdf_kernel|bundle_pred_helper' -s 50 -c 5 -f \
    -o gpurun_out/${tag}_ncu_iter python bench.py --steps 8 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_ncu_iter.log 2>&1
tail -2 gpurun_out/${tag}_ncu_iter.log
ncu -i gpurun_out/${tag}_ncu_iter.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_iter_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_ncu_iter.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${tag}_ncu_iter_src.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_ncu_iter_raw.csv gpurun_out/${tag}_ncu_iter_src.csv > gpurun_out/${tag}_ncu_full_iteration_kernels_bench_config.md 2>&1
python scripts/ncu_traffic.py gpurun_out/${tag}_ncu_iter_raw.csv 1 "bench command, iteration 10 (${tag})" > gpurun_out/${tag}_traffic.log 2>&1
cp profiles/k1_traffic.json gpurun_out/${tag}_k1_traffic.json
# column-pool kernels (K2b) + the person-volume scatter and detail VDF of the final pass, config 5 after 20 iterations
timeout 600 python scripts/cu_step.py 20 3 > gpurun_out/${tag}_cu_step.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:'pool_|sum_kernel' -s 8 -c 4 -f \
    -o gpurun_out/${tag}_ncu_pool python scripts/cu_step.py 20 3 > gpurun_out/${tag}_ncu_pool.log 2>&1
ncu -i gpurun_out/${tag}_ncu_pool.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_pool_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_ncu_pool_raw.csv > gpurun_out/${tag}_ncu_full_column_pool_kernels_bench_config.md 2>&1
rm -f gpurun_out/${tag}_ncu_iter_src.csv
ls -la gpurun_out | grep ${tag}
