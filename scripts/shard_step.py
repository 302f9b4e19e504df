"""What one rank of an N-GPU run does, on one GPU: iterations of the config-5 grid on the origin shard zone % N == 0
(shard-only mode, no collective).  usage: shard_step.py <world_size> [iterations]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlsim_mrm_b200 as D  # noqa: E402
from bench import ensure_workload  # noqa: E402

world = int(sys.argv[1])
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
p = D.read_gmns(ensure_workload("grid100k"), 8)
with D.Engine(p, rank=0, world_size=world, max_columns_per_od=K + 4) as e:
    e.comm_init(None)
    e.iterations(0, 3)
    e.timing(reset=True)
    e.iterations(3, K - 3)   # dtl_iterations: pipelined (DTL_NO_PIPELINE=1: one iteration at a time)
    t = e.timing()
    n = K - 3
    print(f"world {world} kernel {os.environ.get('DTL_SSSP_KERNEL', 'auto')} block {os.environ.get('DTL_SSSP_SMQ_BLOCK', 'auto')}:",
          {k: round(v / n, 3) for k, v in t["ms"].items() if v > 0}, flush=True)
