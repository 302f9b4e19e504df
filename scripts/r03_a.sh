#!/usr/bin/env bash
# round 2 (later session): scatter beside K1 on a second stream -- parity, bench with and without, shards, smq phase profile
mkdir -p gpurun_out
t=r03a
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${t}_pytest.log
tail -3 gpurun_out/${t}_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err; echo "bench rc $?"
DTL_NO_OVERLAP=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${t}_bench_no_overlap.json 2> gpurun_out/${t}_bench_no_overlap.err; echo "bench(no overlap) rc $?"
python - <<'PY'
import json
for f in ("r03a_bench.json", "r03a_bench_no_overlap.json"):
    try:
        b = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, "iter/s", round(b["value"], 2), "e2e", round(b["e2e"]["value"], 2), "K1 ms", round(b["roofline"]["kernel_ms_per_launch"], 3),
              "frac", round(b["roofline"]["frac"], 4), {k: round(v, 3) for k, v in b["phase_ms_per_step"].items()}, "one-stream", round(b.get("one_stream_ms_per_step", 0), 3))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
: > gpurun_out/${t}_shards.log
for world in 2 4 8; do
  for ov in 1 0; do
    ( [ $ov = 0 ] && export DTL_NO_OVERLAP=1
      echo "overlap=$ov $(timeout 200 python scripts/shard_step.py $world 9 2>&1 | tail -1 | cut -c1-260)" ) | tee -a gpurun_out/${t}_shards.log
  done
done
# where a round of the shared-memory-queue kernel goes (thread 0's cycles per phase), 250 contiguous origins
DTL_LIB=$PWD/dlsim-mrm_b200/lib/variants/prof.so PROFILE_CONTIGUOUS=1 timeout 300 python scripts/profile_sssp.py 250 > gpurun_out/${t}_smq_prof.log 2>&1
tail -2 gpurun_out/${t}_smq_prof.log
