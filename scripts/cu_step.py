"""Column-pool updating at the bench configuration: K column-generation iterations, then CU column-pool iterations
(g_update_gradient_cost_and_assigned_flow_in_column_pool: scatter -> VDF -> K2b) with per-phase device times."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlsim_mrm_b200 as D  # noqa: E402
from bench import ensure_workload  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
CU = int(sys.argv[2]) if len(sys.argv) > 2 else 5
p = D.read_gmns(ensure_workload("grid100k"), 8)
with D.Engine(p, max_columns_per_od=K + 4) as e:
    for k in range(K):
        r = e.iteration(k)
    print("last CG iteration", r, flush=True)
    e.timing(reset=True); e.counters(reset=True)
    for n in range(CU):
        r = e.column_update(n)
        print("column update", n, r, flush=True)
    t, c = e.timing(), e.counters()
    print({k: round(v / CU, 3) for k, v in t["ms"].items() if v > 0}, {k: v / CU for k, v in c.items() if v})
    cols = e.columns()
    print("columns", len(cols["n_links"]), "column links", int(cols["n_links"].sum()))
