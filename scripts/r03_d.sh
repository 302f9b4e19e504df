#!/usr/bin/env bash
mkdir -p gpurun_out
t=${1:-r03d}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${t}_pytest.log
tail -3 gpurun_out/${t}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err; echo "bench rc $?"
python - $t <<'PY'
import json, sys
for f in (sys.argv[1] + "_bench.json",):
    try:
        b = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, "iter/s", round(b["value"], 2), "e2e", round(b["e2e"]["value"], 2), "K1 ms", round(b["roofline"]["kernel_ms_per_launch"], 3),
              "frac", round(b["roofline"]["frac"], 4), {k: round(v, 3) for k, v in b["phase_ms_per_step"].items()}, "one-stream", round(b.get("one_stream_ms_per_step", 0), 3),
              {k: round(v["ms_per_step"], 3) for k, v in b["other_kernels"].items()})
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
for world in 2 8; do
  echo "$(timeout 200 python scripts/shard_step.py $world 9 2>&1 | tail -1 | cut -c1-260)" | tee -a gpurun_out/${t}_shards.log
done
