#!/usr/bin/env bash
# multi-GPU bench lines with short timeouts (a hang must not eat the GPU budget): usage r03_scale_safe.sh <tag> <N>...
tag=$1; shift
mkdir -p gpurun_out
export DTL_NCCL_TIMEOUT_S=20
for n in "$@"; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
  echo "bench N=$n rc $?"
  tail -3 gpurun_out/${tag}_bench_${n}gpu.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_bench_${n}gpu.json").read().splitlines() if l.startswith("{")][-1])
    print("N=$n", round(d["value"],2), "iter/s e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["phase_ms_per_step"].items() if v>0}, d["roofline"]["kernel"][:40], d["volume_digest"]["matches_single_gpu_golden"])
except Exception as ex:
    print("N=$n unreadable", ex)
PY
done
