"""Is K1 DRAM/L2-capacity bound?  Same kernel, trees small enough that ~1000 concurrent trees fit in L2."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlsim_mrm_b200 as D
from dlsim_mrm_b200 import synth
for rows, cols, links, zones in ((100, 101, 30000, 1200), (180, 181, 97000, 1200), (316, 317, 300000, 1200)):
    d = tempfile.mkdtemp()
    synth.write_grid(d, rows=rows, cols=cols, n_links=links, n_zones=zones, total_trips=1e5, dest_per_origin=2)
    p = D.read_gmns(d, 8)
    with D.Engine(p) as e:
        cost = p.fftt[0].copy()
        zs = np.arange(1036, dtype=np.int32)
        e.sssp_trees(0, 0, zs, cost=cost, want_trees=False)
        _, _, _, _, st = e.sssp_trees(0, 0, zs, cost=cost, want_trees=False)
        print(f"nodes {p.n_nodes:7d} labels/tree {p.n_nodes*8/1024:7.0f} KB  ms {st['ms']:7.3f}  G relax/s {st['relaxations']/st['ms']/1e6:7.2f}  GTEPS {st['counted_edges']/st['ms']/1e6:7.2f} rounds/tree {st['rounds']/1036:5.0f} us/round {st['ms']*1e3/(st['rounds']/1036):6.2f}", flush=True)
