"""K1 only: time vs number of concurrent origins (latency- or throughput-bound?)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlsim_mrm_b200 as D
from bench import ensure_workload
p = D.read_gmns(ensure_workload("grid100k"), 8)
block = int(os.environ.get("BLOCK", "256"))
with D.Engine(p, block_threads=block) as e:
    cost = p.fftt[0].copy()
    for n in [int(x) for x in sys.argv[1:]] or [37, 74, 148, 296, 592, 888, 1184, 2000]:
        zones = (np.arange(n, dtype=np.int64) * p.n_zones // n).astype(np.int32)
        e.sssp_trees(0, 0, zones, cost=cost, want_trees=False)
        _, _, _, ties, st = e.sssp_trees(0, 0, zones, cost=cost, want_trees=False)
        print(f"block {block} n={n:5d} ms={st['ms']:8.3f} GTEPS={st['counted_edges']/st['ms']/1e6:7.2f} rounds/tree={st['rounds']/n:6.0f} us/round={st['ms']*1e3/(st['rounds']/n):7.2f}", flush=True)
