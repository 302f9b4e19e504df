"""Wall-clock breakdown of the end-to-end sequence bench.py times, with the pipelined call: engine creation from host arrays,
dtl_iterations(0, K), finalize, read-back.  usage: e2e_breakdown2.py [world] [rank] [K]   (world > 1: shard-only mode, no collective)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dlsim_mrm_b200 as D  # noqa: E402
from bench import ensure_workload  # noqa: E402

p = D.read_gmns(ensure_workload("grid100k"), 8)
WORLD = int(sys.argv[1]) if len(sys.argv) > 1 else 1
RANK = int(sys.argv[2]) if len(sys.argv) > 2 else 0
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
for rep in range(3):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    e = D.Engine(p, max_columns_per_od=K + 8, rank=RANK, world_size=WORLD)
    if WORLD > 1:
        e.comm_init(None)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    e.iterations(0, 1)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    e.iterations(1, K - 1)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    e.finalize(K)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    st = e.link_state(0, fields=("tt", "pce_volume"))
    torch.cuda.synchronize(); t.append(time.perf_counter())
    e.close()
    t.append(time.perf_counter())
    d = [round((b - a) * 1e3, 1) for a, b in zip(t, t[1:])]
    print("rep", rep, "create", d[0], "iteration 0", d[1], f"iterations 1..{K - 1}", d[2], "= per iteration", round(d[2] / (K - 1), 2), "finalize", d[3],
          "read-back", d[4], "close", d[5], "total", round(sum(d[:5]), 1), flush=True)
