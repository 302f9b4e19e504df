"""One launch of the batched SSSP kernel on the config-5 grid (for ncu and for sweeps).
usage: profile_sssp.py [n_trees] [block_threads] [delta_factor] [ctas_per_sm]   (PROFILE_CONTIGUOUS=1: the first n zones instead of every (Z/n)-th)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlsim_mrm_b200 as D  # noqa: E402
from bench import ensure_workload  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
block = int(sys.argv[2]) if len(sys.argv) > 2 else 0
delta = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
ctas = int(sys.argv[4]) if len(sys.argv) > 4 else 0
p = D.read_gmns(ensure_workload("grid100k"), 8)
zones = np.arange(n, dtype=np.int32) * (1 if os.environ.get("PROFILE_CONTIGUOUS") else p.n_zones // n)
with D.Engine(p, tree_batch_bytes=n * p.n_nodes * 12 + 1, block_threads=block, delta_factor=delta, ctas_per_sm=ctas) as e:
    cost = p.fftt[0].copy()
    for rep in range(2):
        _, _, _, ties, st = e.sssp_trees(0, 0, zones, cost=cost, want_trees=False)
        print(rep, {k: float(v) for k, v in st.items()}, "GTEPS", round(st["counted_edges"] / st["ms"] / 1e6, 2), flush=True)
