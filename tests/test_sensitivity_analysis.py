"""the sensitivity-analysis / re-routing stage of network_assignment (main_api.cpp:771-1042) -- supply-side
"sa" lane changes (scenario_API.h:152-160) and the re-routing of the agent types with real-time information.
  * CPU: the oracle restatement against the state the UNMODIFIED network_assignment(1, K, CU, 0, S, 0, 4) leaves in the
    reference's globals (tests/golden/sa_ref.npz, made by `make_golden.py sa` through the harness mode golden_dump): link
    travel times and volumes, the whole column pool, bit for bit.
  * GPU: dtl_sa_* against the oracle and the same dumps (columns bit-exact, volumes / times within 1e-6 relative), and the
    output files of the drop-in entry point against the files the reference wrote."""
import gzip
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import FIXTURES, GOLDEN, ROOT, rel_err

CASES = [("sa_corridor", 20, 5, 3), ("sa_corridor", 20, 5, 0), ("corridor_101", 20, 5, 0), ("corridor_101", 20, 5, 2)]
# "dms" rows (information zones, g_column_pool_route_modification, assignment.h:1077-1383): the dataset's own supply_side_scenario.csv
# (one "sa" row that closes a motorway link, one "dms" row); reference dumps from ONE OpenMP thread -- the route modification
# accumulates path travel times on columns that several origins share inside a parallel loop.
DMS_CASES = [("corridor_101+dms", 20, 5, 0), ("corridor_101+dms", 20, 5, 2)]
REL = 1e-6  # link volumes, travel times, column volumes (SURVEY.md 8d parity gates)


@pytest.fixture(scope="module")
def sa():
    return np.load(os.path.join(GOLDEN, "sa_ref.npz"))


def _run(x, p, K, CU, S):
    """the whole entry-point sequence on an oracle or an engine (same method names)"""
    if p.dms_links:
        x.sa_configure(p.rt_info, p.sa_changes, p.dms_links)
    else:
        x.sa_configure(p.rt_info, p.sa_changes)
    for k in range(K):
        x.iteration(k)
    for n in range(CU):
        x.column_update(n)
    x.finalize(K)
    x.sa_begin()
    rows = [x.sa_iteration(it) for it in range(S + 1)]
    if p.dms_links:
        assert x.sa_route_modification() > 0     # main_api.cpp:1036
    cu = [x.sa_column_update(n) for n in range(S)]
    return rows, cu


def _summary_rows(name, S):
    it, cu, stage = [], [], False
    for line in open(os.path.join(GOLDEN, "outputs", f"{name}_sa{S}_final_summary_rows.txt")):
        if line.startswith("Sensitivity analysis stage"):
            stage = True
        elif stage and line[:1].isdigit():
            it.append(float(line.split(",")[1]))
        elif stage and line.startswith("column updating: iteration"):
            cu.append(line.strip())
    return it, cu


@pytest.mark.parametrize("name,K,CU,S", CASES)
def test_reader_scenario_inputs(problems, sa, name, K, CU, S):
    p = problems(name)
    key = f"{name}:S={S}"
    assert np.array_equal(p.rt_info, sa[f"{key}:rt_info"])
    flagged = np.flatnonzero(sa[f"{key}:design_flag"] == -1)
    assert sorted(l for _, l, _ in p.sa_changes) == flagged.tolist()
    for tau, l, lanes in p.sa_changes:
        assert np.float32(lanes) == np.float32(sa[f"{key}:sa_lanes"][l])


def test_reader_keeps_dms_rows(problems):
    p = problems("corridor_101+dms")
    assert p.dms_links == [(0, 395, 98)] and [(t, l) for t, l, _ in p.sa_changes] == [(0, 3898)]   # links 107201->107300, 63702->93601


@pytest.mark.parametrize("name,K,CU,S", CASES + DMS_CASES)
def test_oracle_sa_stage_matches_the_reference(problems, oracle_mod, sa, name, K, CU, S):
    p = problems(name)
    key = f"{name}:S={S}"
    o = oracle_mod.Oracle(p)
    rows, cu = _run(o, p, K, CU, S)
    ls = o.link_state(0)
    assert np.array_equal(ls["tt"], sa[f"{key}:final_tt"])
    assert np.array_equal(ls["pce_volume"], sa[f"{key}:final_pce_volume"])
    assert np.array_equal(ls["person_volume"], sa[f"{key}:final_person_volume"])
    assert np.array_equal(ls["doc"], sa[f"{key}:final_doc"]) and np.array_equal(ls["P"], sa[f"{key}:final_P"])
    oc = o.columns()
    assert np.array_equal(oc["node_sum"], sa[f"{key}:col_node_sum"]) and np.array_equal(oc["seq_no"], sa[f"{key}:col_seq"])
    assert np.array_equal(oc["volume"], sa[f"{key}:col_volume"])
    if f"{key}:col_links" in sa.files:
        assert np.array_equal(oc["links"], sa[f"{key}:col_links"])
    if S > 0:  # path cost / time are refreshed by the SA column-pool iterations
        assert np.array_equal(oc["cost"], sa[f"{key}:col_cost"]) and np.array_equal(oc["time"], sa[f"{key}:col_time"])
    # the rows the reference printed into final_summary.csv (6 significant digits)
    it_ref, cu_ref = _summary_rows(name, S)
    assert len(it_ref) == S + 1 and len(cu_ref) == S
    for r, v in zip(rows, it_ref):
        assert "%g" % r["avg_tt"] == "%g" % v
    for n, (r, line) in enumerate(zip(cu, cu_ref)):
        assert line == (f"column updating: iteration= {n}, avg travel time = {'%g' % r['avg_tt']}(min), optimization obj = "
                        f"{'%g' % r['total_gap']},Relative_gap={'%g' % r['rel_gap_pct']} %")
    # the re-routing did something: columns of the agent types with information carry the SA flag
    if np.any(p.rt_info == 1):
        assert np.count_nonzero(sa[f"{key}:col_sa_flag"]) > 0
    if p.dms_links:  # and the route modification split columns of the message-sign users (impacted_path_flag 3 = diverted, new column)
        assert np.count_nonzero(sa[f"{key}:col_impacted_path_flag"] == 3) > 50 and len(oc["node_sum"]) == len(sa[f"{key}:col_node_sum"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,K,CU,S", CASES + DMS_CASES)
def test_gpu_sa_stage_matches_oracle_and_reference(problems, oracle_mod, sa, name, K, CU, S):
    import dlsim_mrm_b200 as D
    p = problems(name)
    key = f"{name}:S={S}"
    o = oracle_mod.Oracle(p)
    rows_o, cu_o = _run(o, p, K, CU, S)
    with D.Engine(p, max_columns_per_od=K + 4 + S + 1 + (8 if p.dms_links else 0)) as e:
        rows, cu = _run(e, p, K, CU, S)
        for a, b in zip(rows, rows_o):
            assert abs(a["sysTT"] - b["sysTT"]) <= REL * abs(b["sysTT"])
            assert abs(a["avg_tt"] - b["avg_tt"]) <= REL * abs(b["avg_tt"])
        for a, b in zip(cu, cu_o):
            for k2 in ("avg_tt", "total_gap", "rel_gap_pct", "demand"):
                assert abs(a[k2] - b[k2]) <= 1e-5 * max(1e-9, abs(b[k2])), (k2, a, b)
        ls = e.link_state(0)
        assert rel_err(ls["tt"], sa[f"{key}:final_tt"]) <= REL
        assert np.allclose(ls["pce_volume"], sa[f"{key}:final_pce_volume"], rtol=REL, atol=1e-6)
        assert np.allclose(ls["person_volume"], sa[f"{key}:final_person_volume"], rtol=REL, atol=1e-6)
        ec, oc = e.columns(), o.columns()
        assert np.array_equal(ec["node_sum"], sa[f"{key}:col_node_sum"]) and np.array_equal(ec["seq_no"], sa[f"{key}:col_seq"])
        assert np.array_equal(ec["links"], oc["links"])                       # link sequences bit-exact
        assert np.allclose(ec["volume"], sa[f"{key}:col_volume"], rtol=REL, atol=1e-7)
        if S > 0:
            assert np.allclose(ec["time"], sa[f"{key}:col_time"], rtol=REL) and np.allclose(ec["cost"], sa[f"{key}:col_cost"], rtol=REL)
            sc = e.sa_columns()
            assert sc["time"].shape[0] == S
            assert np.array_equal(sc["od_impact"].astype(bool), sa[f"{key}:col_at_od_impacted"].astype(bool))
            assert np.allclose(sc["volume"][-1], sa[f"{key}:col_volume"], rtol=REL, atol=1e-7)


CALL = """
import ctypes, sys
lib = ctypes.CDLL(sys.argv[1])
lib.network_assignment.argtypes = [ctypes.c_int] * 7
lib.network_assignment.restype = ctypes.c_double
assert lib.network_assignment(1, int(sys.argv[2]), int(sys.argv[3]), 0, int(sys.argv[4]), 0, 4) == 1.0
"""


@pytest.mark.gpu
@pytest.mark.parametrize("name,K,CU,S", [("sa_corridor", 20, 5, 3), ("sa_corridor", 20, 5, 0), ("corridor_101", 20, 5, 0),
                                         ("corridor_101+dms", 20, 5, 0)])
def test_output_files_with_sa_stage(native, tmp_path, name, K, CU, S):
    """network_assignment(1, K, CU, 0, S, 0, 4) through the C ABI: link_performance.csv (before / after columns, scenario code),
    route_assignment.csv (re-routed columns, impact flags, per-iteration records) and the summary rows of the stage"""
    from test_output_files_gpu import _compare_csv
    lib = os.path.join(ROOT, "dlsim-mrm_b200", "lib", "libdtalite_b200.so")
    from conftest import stage_fixture
    stage_fixture(name, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", CALL, lib, str(K), str(CU), str(S)], cwd=tmp_path, stdin=subprocess.DEVNULL,
                       capture_output=True, timeout=900)
    assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()
    for f in ("link_performance", "route_assignment"):
        ref = os.path.join(GOLDEN, "outputs", f"{name}_sa{S}_{f}.csv")
        if not os.path.exists(ref):   # large golden files are committed gzip-compressed
            ref = str(tmp_path / f"ref_{f}.csv")
            with gzip.open(os.path.join(GOLDEN, "outputs", f"{name}_sa{S}_{f}.csv.gz"), "rb") as fi, open(ref, "wb") as fo:
                shutil.copyfileobj(fi, fo)
        _compare_csv(tmp_path / f"{f}.csv", ref)
    it_ref, cu_ref = _summary_rows(name, S)
    mine_it, mine_cu, stage = [], [], False
    for line in open(tmp_path / "final_summary.csv"):
        if line.startswith("Sensitivity analysis stage"):
            stage = True
        elif stage and line[:1].isdigit():
            mine_it.append(float(line.split(",")[1]))
        elif stage and line.startswith("column updating: iteration"):
            mine_cu.append(line.strip())
    assert len(mine_it) == S + 1 and len(mine_cu) == S
    for a, b in zip(mine_it, it_ref):
        assert abs(a - b) <= 2e-5 * abs(b)
