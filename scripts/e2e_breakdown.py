"""Wall-clock breakdown of the end-to-end call sequence bench.py times (engine creation from host arrays, iterations,
finalize, read-back) on the config-5 grid."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dlsim_mrm_b200 as D  # noqa: E402
from bench import ensure_workload  # noqa: E402

p = D.read_gmns(ensure_workload("grid100k"), 8)
WORLD = int(sys.argv[1]) if len(sys.argv) > 1 else 1   # > 1: one rank of a sharded run in shard-only mode (no collective)
RANK = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for rep in range(2):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    e = D.Engine(p, max_columns_per_od=24, rank=RANK, world_size=WORLD)
    if WORLD > 1:
        e.comm_init(None)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    for k in range(10):
        e.iteration(k)
        torch.cuda.synchronize(); t.append(time.perf_counter())
    e.finalize(10)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    st = e.link_state(0)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    e.close()
    t.append(time.perf_counter())
    d = [round((b - a) * 1e3, 1) for a, b in zip(t, t[1:])]
    print("rep", rep, "create", d[0], "iterations", d[1:11], "finalize", d[11], "read-back", d[12], "close", d[13], "total", round(sum(d[:13]), 1), flush=True)
