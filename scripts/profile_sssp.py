"""One launch of the batched SSSP kernel on the config-5 grid for ncu (short: the profiler replays it ~40 times)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlsim_mrm_b200 as D  # noqa: E402
from bench import ensure_workload  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
p = D.read_gmns(ensure_workload("grid100k"), 8)
zones = np.arange(n, dtype=np.int32) * (p.n_zones // n)
with D.Engine(p, tree_batch_bytes=n * p.n_nodes * 12 + 1) as e:
    cost = p.fftt[0].copy()
    for rep in range(2):
        _, _, _, ties, st = e.sssp_trees(0, 0, zones, cost=cost, want_trees=False)
        print(rep, st, "GTEPS", st["counted_edges"] / st["ms"] / 1e6, flush=True)
