#!/bin/bash
# one gpurun call: GPU parity tests on the main library, then A/B of kernel variants on the config-5 grid
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
shift 0
bash scripts/ab.sh "$@" 2>&1 | tee gpurun_out/ab.log
