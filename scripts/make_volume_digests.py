"""SHA-256 of the Q31.32 link-volume vector of the config-5 grid after n = 1..N assignment iterations on ONE GPU ->
tests/golden/grid100k_volume_digests.json.  bench.py compares the vector every rank of a multi-GPU run holds after its
W + K iterations with this table: the sharded run must reproduce the single-GPU integers bit for bit.
usage (GPU box): python scripts/make_volume_digests.py [N=40]"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlsim_mrm_b200 as D  # noqa: E402
from bench import ensure_workload  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "tests", "golden", "grid100k_volume_digests.json")
p = D.read_gmns(ensure_workload("grid100k"), 8)
table = {}
with D.Engine(p, max_columns_per_od=N + 4) as e:
    for k in range(N):
        e.iteration(k)
        table[str(k + 1)] = hashlib.sha256(e.volume_fixed(0).tobytes()).hexdigest()
with open(out_path, "w") as f:
    json.dump({"what": "sha256 of dtl_get_volume_fixed(tau 0) after n iterations of grid100k on one GPU (max_columns_per_od = n + 4 or more)",
               "digests": table}, f, indent=1)
print("wrote", out_path, len(table))
