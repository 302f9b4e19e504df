"""The file contract of the drop-in (SURVEY.md 8b "Outputs"): link_performance.csv and route_assignment.csv written by
the B200 executable against the files the reference's own network_assignment() writes on the same inputs
(tests/golden/outputs/, made by tests/golden/make_golden.py).  Same header, same rows in the same order, text fields
identical, numeric fields within the path's tolerance at the precision the reference prints."""
import csv
import os
import shutil
import subprocess

import pytest

from conftest import FIXTURES, GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _num(x):
    try:
        return float(x)
    except ValueError:
        return None


def _compare_csv(mine_path, ref_path, rel=1e-6, report=20):
    with open(mine_path, newline="") as f:
        mine = list(csv.reader(f))
    with open(ref_path, newline="") as f:
        ref = list(csv.reader(f))
    assert mine[0] == ref[0], f"header differs:\n{mine[0]}\n{ref[0]}"
    problems = []
    if len(mine) != len(ref):
        problems.append(f"{len(mine)} rows instead of {len(ref)}")
    for r, (a, b) in enumerate(zip(mine[1:], ref[1:]), 2):
        if len(a) != len(b):
            problems.append(f"row {r}: {len(a)} fields instead of {len(b)}")
            continue
        for c, (x, y) in enumerate(zip(a, b)):
            if x == y:
                continue
            fx, fy = _num(x), _num(y)
            col = ref[0][c] if c < len(ref[0]) else f"#{c}"
            if fx is None or fy is None:
                # a coordinate list or a ;-joined sequence: compare token by token
                tx, ty = x.replace(",", " ").split(), y.replace(",", " ").split()
                if len(tx) != len(ty) or any(p != q and (_num(p) is None or _num(q) is None or abs(_num(p) - _num(q)) > 1e-5)
                                             for p, q in zip(tx, ty)):
                    problems.append(f"row {r} {col}: {x[:60]!r} instead of {y[:60]!r}")
                continue
            # the reference prints 1 to 6 decimals: allow one unit of the printed last place on top of the relative tolerance
            dec = len(y.split(".")[1]) if "." in y else 0
            if abs(fx - fy) > rel * max(abs(fx), abs(fy)) + 1.01 * 10.0 ** (-dec):
                problems.append(f"row {r} {col}: {x} instead of {y}")
    assert not problems, f"{len(problems)} differences, first {report}:\n" + "\n".join(problems[:report])


CALL = """
import ctypes, sys
lib = ctypes.CDLL(sys.argv[1])
lib.network_assignment.argtypes = [ctypes.c_int] * 7
lib.network_assignment.restype = ctypes.c_double
assert lib.network_assignment(1, int(sys.argv[2]), int(sys.argv[3]), 0, 0, 0, 4) == 1.0
"""


@pytest.mark.parametrize("name,K,CU", [("two_corridor", 20, 0), ("corridor_3", 20, 20), ("edge_cases", 10, 4), ("corridor_3_2p", 12, 6),
                                       ("toll_signal", 10, 4), ("asu_qvdf", 20, 3)])
def test_output_files_match_the_reference(native, tmp_path, name, K, CU):
    """network_assignment(1, K, CU, 0, 0, 0, 4) through the C ABI, in the dataset directory, like a ctypes caller of the
    reference's library; the golden files come from the same call on the reference."""
    import sys
    lib = os.path.join(ROOT, "dlsim-mrm_b200", "lib", "libdtalite_b200.so")
    src = os.path.join(FIXTURES, name)
    for f in os.listdir(src):
        shutil.copy(os.path.join(src, f), tmp_path)
    r = subprocess.run([sys.executable, "-c", CALL, lib, str(K), str(CU)], cwd=tmp_path, stdin=subprocess.DEVNULL,
                       capture_output=True, timeout=600)
    assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()
    for f in ("link_performance", "route_assignment"):
        _compare_csv(tmp_path / f"{f}.csv", os.path.join(GOLDEN, "outputs", f"{name}_{f}.csv"))
    # final_summary.csv: iteration rows (iteration, CPU time, agents, average travel time, gap %) and column-pool rows
    import re
    def pick(path, after_header):
        rows, seen = [], not after_header
        for l in open(path):
            seen = seen or l.startswith("Iteration, CPU Running Time")
            if seen and (l[:1].isdigit() and l.count(",") == 4 or l.startswith("column updating: iteration")):
                rows.append(l.strip())
        return rows
    mine, ref = pick(tmp_path / "final_summary.csv", True), pick(os.path.join(GOLDEN, "outputs", f"{name}_final_summary_rows.txt"), False)
    assert len(mine) == len(ref) == K + CU, (len(mine), len(ref))
    num = re.compile(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?")
    for a_, b_ in zip(mine, ref):
        if a_[:1].isdigit():
            x, y = a_.split(","), b_.split(",")
            assert x[0] == y[0] and x[2] == y[2], (a_, b_)              # iteration, number of agents; CPU time is not compared
            for u, v in zip(x[3:5], y[3:5]):                             # 6 significant digits are printed
                assert abs(float(u) - float(v)) <= 2e-5 * max(1e-3, abs(float(v))), (a_, b_)
        else:
            assert num.sub("#", a_) == num.sub("#", b_), (a_, b_)       # same text around the numbers
            for u, v in zip(num.findall(a_), num.findall(b_)):
                assert abs(float(u) - float(v)) <= 2e-5 * max(1e-3, abs(float(v))), (a_, b_)
