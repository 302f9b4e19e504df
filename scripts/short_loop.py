"""A short assignment loop on the config-5 grid for ncu: 5 column-generation iterations + 1 column-pool update on a
500-origin-zone shard (rank 0 of 4), so that one --set full pass over each kernel stays short."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlsim_mrm_b200 as D  # noqa: E402
from bench import ensure_workload  # noqa: E402

p = D.read_gmns(ensure_workload("grid100k"), 8)
with D.Engine(p, rank=0, world_size=4, max_columns_per_od=12) as e:
    e.comm_init(None)  # shard-only mode: no collective
    for k in range(5):
        print(k, e.iteration(k), flush=True)
    print(e.column_update(0), flush=True)
    e.finalize(5)
    print(e.timing())
