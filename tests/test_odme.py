"""OD demand estimation against link counts -- Assignment::Demand_ODME (ODME.h:91-516) and its scatter
g_reset_and_update_link_volume_based_on_ODME_columns (assignment.h:263-563).
  * CPU: the oracle restatement against the state the UNMODIFIED network_assignment(1, K, CU, ODME, 0, 0, 4) leaves in the
    reference's globals (tests/golden/odme_ref.npz, `make_golden.py odme`, one OpenMP thread: the reference sums link volumes
    inside a parallel loop) and against the "ODME #s" rows it printed: bit for bit / digit for digit.
  * GPU: dtl_odme_* against the oracle and the same dumps (column volumes, link volumes, travel times within 1e-6 relative;
    the stopping iteration identical), and the output files of the drop-in entry point."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, rel_err, stage_fixture

CASES = [("sa_corridor+odme", 20, 5, 12), ("corridor_101+odme", 20, 5, 20)]
REL = 1e-6


@pytest.fixture(scope="module")
def od():
    return np.load(os.path.join(GOLDEN, "odme_ref.npz"))


def _fmt_row(s, r):
    """the reference's own ostream formatting of a summary row (assignment.h:539-546)"""
    return ("ODME #%d, link MAE= %g,link_MAPE: %g%%,system_MPE: %g%%,avg_tt = %g(min) ,UE gap =%g(min) = (%g %%)"
            % (s, r["link_mae"], np.float32(r["gap"]) * np.float32(100), np.float32(r["system_gap"]) * np.float32(100),  # float products
               r["avg_tt"], r["ue_gap"], r["ue_gap_pct"]))


def _run(x, p, K, CU, n):
    """assignment, ODME loop with the reference's stopping rule (ODME.h:240-260), closing scatter, SA stage with 0 iterations"""
    x.sa_configure(p.rt_info, p.sa_changes)
    for k in range(K):
        x.iteration(k)
    for c in range(CU):
        x.column_update(c)
    x.finalize(K)
    x.odme_configure(p.at_pce, p.at_occ, p.sensors)
    rows, prev, stopped = [], 9999999.0, None
    for s in range(n):
        r = x.odme_scatter(s)
        rows.append((s, r))
        if s >= 5 and r["gap"] - prev > 0.01:
            stopped = s
            break
        prev = r["gap"]
        x.odme_update()
    rows.append((n, x.odme_scatter(n)))
    x.sa_begin()
    x.sa_iteration(0)
    return rows, stopped


def _ref_rows(name, n):
    rows, stop = [], None
    for line in open(os.path.join(GOLDEN, "outputs", f"{name}_odme{n}_final_summary_rows.txt")):
        if line.startswith("ODME #"):
            rows.append(line.rstrip("\n"))
        if line.startswith("ODME stage terminates"):
            stop = int(line.split("at iteration =")[1])
    return rows, stop


@pytest.mark.parametrize("name,K,CU,n", CASES)
def test_reader_measurements(problems, od, name, K, CU, n):
    p = problems(name)
    key = f"{name}:ODME={n}"
    obs = np.zeros(p.n_links); ub = np.ones(p.n_links, np.int32)
    for tau, l, c, u in p.sensors:                      # the overwrite rule of ODME.h:156-175
        if obs[l] >= 1:
            if u == 0:
                obs[l], ub[l] = c, u
        else:
            obs[l], ub[l] = c, u
    assert np.array_equal(obs, od[f"{key}:obs_count"]) and np.array_equal(ub, od[f"{key}:upper_bound_flag"])
    assert np.array_equal(p.at_pce.astype(np.float64), od[f"{key}:agent_pce"]) and np.array_equal(p.at_occ.astype(np.float64), od[f"{key}:agent_occ"])


@pytest.mark.parametrize("name,K,CU,n", CASES)
def test_oracle_odme_matches_the_reference(problems, oracle_mod, od, name, K, CU, n):
    p = problems(name)
    key = f"{name}:ODME={n}"
    o = oracle_mod.Oracle(p)
    rows, stopped = _run(o, p, K, CU, n)
    ref_rows, ref_stop = _ref_rows(name, n)
    assert stopped == ref_stop
    assert [_fmt_row(s, r) for s, r in rows] == ref_rows          # every printed digit
    ls = o.link_state(0)
    assert np.array_equal(ls["tt"], od[f"{key}:final_tt"]) and np.array_equal(ls["pce_volume"], od[f"{key}:final_pce_volume"])
    assert np.array_equal(ls["person_volume"], od[f"{key}:final_person_volume"])
    oc = o.columns()
    assert np.array_equal(oc["node_sum"], od[f"{key}:col_node_sum"]) and np.array_equal(oc["volume"], od[f"{key}:col_volume"])
    assert np.array_equal(oc["cost"], od[f"{key}:col_cost"])      # the path gradient of the last step (ODME.h:426)


@pytest.mark.gpu
@pytest.mark.parametrize("name,K,CU,n", CASES)
def test_gpu_odme_matches_oracle_and_reference(problems, oracle_mod, od, name, K, CU, n):
    import dlsim_mrm_b200 as D
    p = problems(name)
    key = f"{name}:ODME={n}"
    o = oracle_mod.Oracle(p)
    rows_o, stop_o = _run(o, p, K, CU, n)
    with D.Engine(p, max_columns_per_od=K + 6) as e:
        rows, stopped = _run(e, p, K, CU, n)
        assert stopped == stop_o == _ref_rows(name, n)[1]
        for (s, a), (_, b) in zip(rows, rows_o):
            for k2 in ("gap", "system_gap", "avg_tt", "ue_gap", "ue_gap_pct", "demand"):
                assert abs(a[k2] - b[k2]) <= 1e-5 * max(1e-9, abs(b[k2])), (s, k2, a, b)
            assert abs(a["link_mae"] - b["link_mae"]) <= 1.0 / max(1, a["links_with_data"]) + 1e-5 * abs(b["link_mae"])  # integer abs() per link
            assert a["links_with_data"] == b["links_with_data"]
        ls = e.link_state(0)
        assert rel_err(ls["tt"], od[f"{key}:final_tt"]) <= REL
        assert np.allclose(ls["pce_volume"], od[f"{key}:final_pce_volume"], rtol=REL, atol=1e-6)
        assert np.allclose(ls["person_volume"], od[f"{key}:final_person_volume"], rtol=REL, atol=1e-6)
        ec = e.columns()
        assert np.array_equal(ec["node_sum"], od[f"{key}:col_node_sum"])
        assert np.allclose(ec["volume"], od[f"{key}:col_volume"], rtol=REL, atol=1e-7)
        st = e.odme_state(0)
        assert np.array_equal(st["obs_count"], od[f"{key}:obs_count"]) and np.array_equal(st["upper_bound_flag"], od[f"{key}:upper_bound_flag"])
        m = od[f"{key}:obs_count"] >= 1
        assert np.allclose(st["est_count_dev"][m], od[f"{key}:est_count_dev"][m], rtol=1e-5, atol=1e-4)


@pytest.mark.gpu
def test_gpu_odme_run_wrapper(problems, oracle_mod):
    """dtl_odme_run = the same loop in one call"""
    import dlsim_mrm_b200 as D
    p = problems("sa_corridor+odme")
    with D.Engine(p) as e:
        for k in range(6):
            e.iteration(k)
        e.finalize(6)
        e.odme_configure(p.at_pce, p.at_occ, p.sensors)
        v0 = e.columns()["volume"].copy()
        rows, done = e.odme_run(8)
        assert done == 8 and rows.shape == (9, 8) and rows[-1, 7] == 5
        assert rows[-1, 0] < rows[0, 0]                       # the link MAPE went down
        assert not np.array_equal(e.columns()["volume"], v0)


CALL = """
import ctypes, sys
lib = ctypes.CDLL(sys.argv[1])
lib.network_assignment.argtypes = [ctypes.c_int] * 7
lib.network_assignment.restype = ctypes.c_double
assert lib.network_assignment(1, int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), 0, 0, 4) == 1.0
"""


@pytest.mark.gpu
def test_output_files_with_odme(native, tmp_path):
    """network_assignment(1, 20, 5, 12, 0, 0, 4) through the C ABI: the count columns of link_performance.csv, the estimated
    path volumes of route_assignment.csv and the ODME rows of final_summary.csv against the reference's own files"""
    from test_output_files_gpu import _compare_csv
    name, K, CU, n = "sa_corridor+odme", 20, 5, 12
    lib = os.path.join(ROOT, "dlsim-mrm_b200", "lib", "libdtalite_b200.so")
    stage_fixture(name, tmp_path)
    r = subprocess.run([sys.executable, "-c", CALL, lib, str(K), str(CU), str(n)], cwd=tmp_path, stdin=subprocess.DEVNULL,
                       capture_output=True, timeout=900)
    assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()
    for f in ("link_performance", "route_assignment"):
        _compare_csv(tmp_path / f"{f}.csv", os.path.join(GOLDEN, "outputs", f"{name}_odme{n}_{f}.csv"))
    mine = [l.rstrip("\n") for l in open(tmp_path / "final_summary.csv") if l.startswith("ODME")]
    ref = [l.rstrip("\n") for l in open(os.path.join(GOLDEN, "outputs", f"{name}_odme{n}_final_summary_rows.txt"))]
    assert len(mine) == len(ref)
    import re
    num = re.compile(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?")
    for a, b in zip(mine, ref):
        assert num.sub("#", a) == num.sub("#", b), (a, b)
        for u, v in zip(num.findall(a), num.findall(b)):
            assert abs(float(u) - float(v)) <= 2e-4 * max(1.0, abs(float(v))), (a, b)
