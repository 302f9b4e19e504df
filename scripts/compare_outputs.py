"""Run network_assignment through the C ABI on the fixtures and summarise the differences between the files written and
the reference's own files (tests/golden/outputs): per column, how many cells differ and one example."""
import collections
import csv
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_output_files_gpu import CALL, _num  # noqa: E402

for name, K, CU in (("two_corridor", 20, 0), ("corridor_3", 20, 20), ("edge_cases", 10, 4), ("corridor_3_2p", 12, 6)):
    tmp = tempfile.mkdtemp()
    src = os.path.join(ROOT, "tests", "fixtures", name)
    for f in os.listdir(src):
        shutil.copy(os.path.join(src, f), tmp)
    r = subprocess.run([sys.executable, "-c", CALL, os.path.join(ROOT, "dlsim-mrm_b200", "lib", "libdtalite_b200.so"), str(K), str(CU)],
                       cwd=tmp, stdin=subprocess.DEVNULL, capture_output=True)
    if r.returncode:
        print(name, "FAILED", r.stderr.decode()[-500:])
        continue
    for f in ("link_performance", "route_assignment"):
        mine = list(csv.reader(open(os.path.join(tmp, f + ".csv"), newline="")))
        ref = list(csv.reader(open(os.path.join(ROOT, "tests", "golden", "outputs", f"{name}_{f}.csv"), newline="")))
        print(f"== {name} {f}: rows {len(mine)} vs {len(ref)}; header equal {mine[0] == ref[0]}")
        cnt, ex = collections.Counter(), {}
        for rno, (a, b) in enumerate(zip(mine[1:], ref[1:]), 2):
            if len(a) != len(b):
                cnt["<fields>"] += 1
                ex.setdefault("<fields>", (rno, len(a), len(b)))
                continue
            for c, (x, y) in enumerate(zip(a, b)):
                if x == y:
                    continue
                fx, fy = _num(x), _num(y)
                dec = len(y.split(".")[1]) if "." in y else 0
                if fx is not None and fy is not None and abs(fx - fy) <= 1e-6 * max(abs(fx), abs(fy)) + 1.01 * 10.0 ** (-dec):
                    continue
                col = ref[0][c] if c < len(ref[0]) else f"#{c}"
                cnt[col] += 1
                ex.setdefault(col, (rno, x[:70], y[:70]))
        for col, n in cnt.most_common():
            print(f"   {col}: {n} cells, e.g. row {ex[col][0]}: mine {ex[col][1]!r} ref {ex[col][2]!r}")
    shutil.rmtree(tmp, ignore_errors=True)
