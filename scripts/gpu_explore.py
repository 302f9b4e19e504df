"""Scratch exploration on the GPU box: first timings of the config-5 grid and a parameter sweep of K1."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlsim_mrm_b200 as D  # noqa: E402
from dlsim_mrm_b200 import synth  # noqa: E402

out = {}
work = sys.argv[1] if len(sys.argv) > 1 else "grid100k"
d = f"/tmp/{work}"
t = time.time(); synth.write_grid(d, **synth.WORKLOADS[work]); out["gen_s"] = time.time() - t
t = time.time(); p = D.read_gmns(d, 8); out["read_s"] = time.time() - t
print("problem", p.n_nodes, p.n_links, p.n_zones, p.n_od, flush=True)
zones = np.arange(0, p.n_zones, max(1, p.n_zones // 296), dtype=np.int32)[:296]
sweep = []
for block in (128, 256, 512):
    for ctas in (0,):
        for delta in (2.0, 4.0, 8.0, 16.0, 32.0):
            with D.Engine(p, block_threads=block, ctas_per_sm=ctas, delta_factor=delta) as e:
                e.vdf(0, np.zeros(p.n_links))      # free-flow costs into the forward star
                cost = None
                # costs: engine VDF with write_cost=False in dtl_vdf; use fftt-based host costs instead
                c = p.fftt[0].copy()
                _, _, _, ties, st = e.sssp_trees(0, 0, zones, cost=c, want_trees=False)
                _, _, _, ties, st = e.sssp_trees(0, 0, zones, cost=c, want_trees=False)
                row = dict(block=block, ctas=ctas, delta=delta, ms=st["ms"], gteps=st["counted_edges"] / st["ms"] / 1e6,
                           relax_ratio=st["relaxations"] / max(1, st["counted_edges"]), rounds_per_tree=st["rounds"] / zones.size,
                           pops_per_tree=st["pops"] / zones.size)
                print(row, flush=True)
                sweep.append(row)
out["sweep"] = sweep
with D.Engine(p) as e:
    print("device bytes", e.device_bytes / 2**30, "GiB", "tasks", e.n_tasks, flush=True)
    rows = []
    for k in range(4):
        t = time.time(); r = e.iteration(k); dt = time.time() - t
        tm = e.timing(reset=True); c = e.counters(reset=True)
        rows.append(dict(k=k, wall_s=dt, row=r, ms=tm["ms"], counters=c))
        print(rows[-1], flush=True)
    out["iterations"] = rows
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"explore_{work}.json"), "w"), indent=1, default=float)
