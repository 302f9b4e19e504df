#!/usr/bin/env bash
mkdir -p gpurun_out
t=${1:-r03p}
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k "pipelined or iterations_call" > gpurun_out/${t}_pytest_pipe.log 2>&1; echo "pytest(pipelined) rc $?"; tail -15 gpurun_out/${t}_pytest_pipe.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err; echo "bench rc $?"; tail -3 gpurun_out/${t}_bench.err
DTL_NO_PIPELINE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${t}_bench_nopipe.json 2> gpurun_out/${t}_bench_nopipe.err; echo "bench(no pipeline) rc $?"
python - $t <<'PY'
import json, sys
for f in (sys.argv[1] + "_bench.json", sys.argv[1] + "_bench_nopipe.json"):
    try:
        b = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, "iter/s", round(b["value"], 2), "e2e", round(b["e2e"]["value"], 2), "K1 ms", round(b["roofline"]["kernel_ms_per_launch"], 3),
              "frac", round(b["roofline"]["frac"], 4), {k: round(v, 3) for k, v in b["phase_ms_per_step"].items()}, "wall/step", round(b["wall_ms_per_step"], 3))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
