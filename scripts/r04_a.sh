#!/usr/bin/env bash
# round 2, last evidence run: the destination-bounded search is the engine's default.  GPU parity tests, smoke(), the bench line;
# then (only when all of them passed) the launch list and one ncu --set full of the K1 launch for roofline.traffic.   tag = $1
tag=${1:-r04a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; prc=$?; echo "pytest rc $prc" | tee -a gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err; rc=$?; echo "bench rc $rc"
tail -5 gpurun_out/${tag}_bench_1gpu.err | cut -c1-400
[ $rc -ne 0 ] && exit 1
[ $prc -ne 0 ] && exit 1
[ -n "$NO_NCU" ] && exit 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_ncu_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'sssp_bundle_async_kernel' -s 10 -c 1 -f \
    -o gpurun_out/${tag}_ncu_k1 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_ncu_k1.log 2>&1
tail -2 gpurun_out/${tag}_ncu_k1.log
ncu -i gpurun_out/${tag}_ncu_k1.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_k1_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_ncu_k1_raw.csv > gpurun_out/${tag}_ncu_full_k1_bounded_bench_config.md 2>&1
python scripts/ncu_traffic.py gpurun_out/${tag}_ncu_k1_raw.csv 1 "bench command, iteration 10 (${tag}); destination-bounded search (the default)" > gpurun_out/${tag}_traffic.log 2>&1
cp profiles/k1_traffic.json gpurun_out/${tag}_k1_traffic.json
ls -la gpurun_out | grep ${tag}
