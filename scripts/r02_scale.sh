#!/usr/bin/env bash
# multi-GPU evidence on one box: NCCL parity tests (config 4 over 2/4/8 ranks), then the bench line at N = 8, 4, 2 (torchrun, one rank per GPU)
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
timeout 900 python -m pytest tests/test_nccl_gpu.py -q -m gpu > gpurun_out/${tag}_nccl_pytest.log 2>&1; echo "pytest rc $?" | tee -a gpurun_out/${tag}_nccl_pytest.log
tail -3 gpurun_out/${tag}_nccl_pytest.log
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
  echo "bench N=$n rc $?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_bench_${n}gpu.json").read().splitlines() if l.startswith("{")][-1])
    print("N=$n", round(d["value"],2), "iter/s e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["phase_ms_per_step"].items() if v>0}, d["roofline"]["kernel"][:50], d["volume_digest"])
except Exception as ex:
    print("N=$n unreadable", ex)
PY
done
