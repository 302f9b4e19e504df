#!/usr/bin/env bash
# round 2, first GPU call: parity tests, then the bench line with and without the node-major tree path
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc $?"
DTL_NO_SLIM_TREES=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_noslim.json 2> gpurun_out/r02a_bench_noslim.err; echo "bench(noslim) rc $?"
python - <<'PY'
import json
for f in ("r02a_bench.json", "r02a_bench_noslim.json"):
    try:
        b = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, "iter/s", round(b["value"], 2), "e2e", round(b["e2e"]["value"], 2), "K1 ms", round(b["roofline"]["kernel_ms_per_launch"], 3),
              "frac", round(b["roofline"]["frac"], 4), {k: round(v, 3) for k, v in b["phase_ms_per_step"].items()})
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
