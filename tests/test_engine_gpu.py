"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same inputs and
against the golden state dumped from the reference engine.  Bars (BASELINE.json north_star):
  * shortest-path labels, predecessor trees and path columns bit-exact (ties by the reference's rule),
  * link volumes, travel times and relative gap within 1e-6 relative.
"""
import hashlib
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import FIXTURES, GOLDEN, ROOT, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-6  # north_star tolerance for volumes / travel times / gap


@pytest.fixture(scope="module")
def D():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import dlsim_mrm_b200 as D
    return D


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------------------------------------ K1
# global-memory queues (many trees per SM) / shared-memory queues (few trees per SM) / origin bundles (B origins per search;
# "bundle": bundle size by tree count = round-synchronous bundles of 4 at these sizes, "bundle16": bundles of 16 = the search
# without round barriers that the bench configuration runs)
KERNELS = ["gq", "smq", "bundle", "bundle16"]


def _force_kernel(monkeypatch, kernel):
    monkeypatch.setenv("DTL_SSSP_KERNEL", "bundle" if kernel.startswith("bundle") else kernel)
    if kernel == "bundle16":
        monkeypatch.setenv("DTL_BUNDLE_B", "16")


@pytest.mark.parametrize("kernel", KERNELS)
def test_sssp_known_answer_corridor_101(D, problems, golden, oracle_mod, monkeypatch, kernel):
    """K1 fed the reference's own generalized costs: labels, pred nodes and pred links bit-exact for every origin."""
    _force_kernel(monkeypatch, kernel)
    p, g = problems("corridor_101"), golden("corridor_101")
    o = oracle_mod.Oracle(p)
    with D.Engine(p) as e:
        for k, at in ((0, 0), (2, 0), (19, 2)):
            cost = g[f"gen_cost_k{k}_at{at}"]
            zones = np.arange(p.n_zones, dtype=np.int32)
            lab, pn, pl, ties, st = e.sssp_trees(at, 0, zones, cost=cost)
            fs = o.forward_star(at, 0)
            e_from = np.repeat(np.arange(p.n_nodes), np.diff(fs[0]))
            e_tag = p.link_tag[fs[1]]
            counted = 0
            for z in zones:
                rl, rpn, rpl, pops, relax = o.sssp(fs, cost, int(z))
                assert np.array_equal(lab[z].view(np.int64), rl.view(np.int64)), f"labels differ, k={k} zone {z}"
                assert np.array_equal(pl[z], rpl), f"pred links differ, k={k} zone {z}"
                assert np.array_equal(pn[z], rpn), f"pred nodes differ, k={k} zone {z}"
                counted += int(np.count_nonzero((rl < 1e15)[e_from] & ((e_tag < 0) | (e_tag == z))))
            assert st["counted_edges"] == counted
            assert st["relaxations"] >= counted
            # golden full trees of zone 7 (reference engine itself)
            tagk = f"k={k}:at={at}:tau=0:z=7"
            assert np.array_equal(lab[7], g["tree_label:" + tagk]) and np.array_equal(pl[7], g["tree_pred_link:" + tagk])
            assert np.array_equal(pn[7], g["tree_pred_node:" + tagk])


@pytest.mark.parametrize("B,block,cluster", [(4, 1024, 1), (8, 512, 1), (16, 1024, 1), (16, 512, 0), (8, 1024, 0)])
def test_sssp_bundle_sizes(D, problems, golden, oracle_mod, monkeypatch, B, block, cluster):
    """Round-synchronous origin-bundle kernel at every bundle size / CTA size, with and without spatial clustering of the origins."""
    monkeypatch.setenv("DTL_SSSP_KERNEL", "bundle")
    monkeypatch.setenv("DTL_BUNDLE_ASYNC", "0")
    monkeypatch.setenv("DTL_BUNDLE_B", str(B))
    monkeypatch.setenv("DTL_BUNDLE_BLOCK", str(block))
    if not cluster:
        monkeypatch.setenv("DTL_BUNDLE_NO_CLUSTER", "1")
    p, g = problems("corridor_101"), golden("corridor_101")
    o = oracle_mod.Oracle(p)
    with D.Engine(p) as e:
        for k, at in ((2, 0), (19, 2)):
            cost = g[f"gen_cost_k{k}_at{at}"]
            zones = np.arange(p.n_zones, dtype=np.int32)[::-1].copy()   # 99 origins: the last bundle is ragged
            lab, pn, pl, ties, st = e.sssp_trees(at, 0, zones, cost=cost)
            fs = o.forward_star(at, 0)
            for i, z in enumerate(zones):
                rl, rpn, rpl, pops, relax = o.sssp(fs, cost, int(z))
                assert np.array_equal(lab[i].view(np.int64), rl.view(np.int64)), f"labels differ, k={k} zone {z}"
                assert np.array_equal(pl[i], rpl) and np.array_equal(pn[i], rpn), f"predecessors differ, k={k} zone {z}"
            assert st["replayed"] == 0 and e.counters()["reruns"] == 0


@pytest.mark.parametrize("B", [16, 8])
def test_sssp_bundle_async_window(D, problems, golden, oracle_mod, monkeypatch, B):
    """Origin-bundle search without round barriers inside a label window (ring-buffer near queue; the default for bundles of
    8 and 16 origins): same bits."""
    monkeypatch.setenv("DTL_SSSP_KERNEL", "bundle")
    monkeypatch.delenv("DTL_BUNDLE_ASYNC", raising=False)
    monkeypatch.setenv("DTL_BUNDLE_B", str(B))
    p, g = problems("corridor_101"), golden("corridor_101")
    o = oracle_mod.Oracle(p)
    with D.Engine(p) as e:
        for k, at in ((2, 0), (19, 2)):
            cost = g[f"gen_cost_k{k}_at{at}"]
            zones = np.arange(p.n_zones, dtype=np.int32)
            lab, pn, pl, ties, st = e.sssp_trees(at, 0, zones, cost=cost)
            fs = o.forward_star(at, 0)
            for i, z in enumerate(zones):
                rl, rpn, rpl, pops, relax = o.sssp(fs, cost, int(z))
                assert np.array_equal(lab[i].view(np.int64), rl.view(np.int64)), f"labels differ, k={k} zone {z}"
                assert np.array_equal(pl[i], rpl) and np.array_equal(pn[i], rpn), f"predecessors differ, k={k} zone {z}"
            assert st["replayed"] == 0 and e.counters()["reruns"] == 0


def _tie_grid(n=12):
    """Uniform n x n grid with unit costs: every interior node has two tight predecessors."""
    import dlsim_mrm_b200 as D
    N = n * n
    links = []
    for r in range(n):
        for c in range(n):
            u = r * n + c
            if c + 1 < n: links += [(u, u + 1), (u + 1, u)]
            if r + 1 < n: links += [(u, u + n), (u + n, u)]
    zones = [0, N - 1, n - 1]
    cent = [N + i for i in range(len(zones))]
    lf = [a for a, b in links]; lt = [b for a, b in links]; tag = [-1] * len(links)
    for zi, node in enumerate(zones):
        lf += [node, cent[zi]]; lt += [cent[zi], node]; tag += [-1, zi]
    L, Z = len(lf), len(zones)
    p = D.Problem(N + Z, L, Z, 1, 1, 4, 6.0)
    nz = np.full(N + Z, -1, np.int32); nz[cent] = np.arange(Z)
    p.arrays.update(link_from=np.array(lf, np.int32), link_to=np.array(lt, np.int32), link_tag=np.array(tag, np.int32),
                    node_zone=nz, zone_centroid=np.array(cent, np.int32),
                    fftt=np.ones((1, L)), cap=np.full((1, L), 1000.0), vf=np.full((1, L), 60.0), vc=np.full((1, L), 51.0),
                    nlanes=np.ones((1, L)), cap_lane=np.full((1, L), 1000.0), qdf=np.ones((1, L)), t2=np.full((1, L), 7.5),
                    L_h=np.ones((1, L)), start_h=np.full((1, L), 7.0), end_h=np.full((1, L), 8.0), vot=np.array([10.0], np.float32),
                    od_o=np.array([0, 0, 1], np.int32), od_d=np.array([1, 2, 0], np.int32), od_at=np.zeros(3, np.int32),
                    od_tau=np.zeros(3, np.int32), od_vol=np.array([1.0, 2.0, 3.0]), origin_demand=np.array([3, 3, 0], np.float32))
    return p.with_defaults().finalize()


@pytest.mark.parametrize("kernel", KERNELS)
def test_sssp_ties_follow_the_reference_rule(D, oracle_mod, monkeypatch, kernel):
    """Exact FP64 ties: the engine detects them and replays the reference's deque, so the trees still match."""
    _force_kernel(monkeypatch, kernel)
    p = _tie_grid()
    o = oracle_mod.Oracle(p)
    cost = np.ones(p.n_links)
    cost[p.link_tag >= 0] = 0.0
    cost[(p.link_to >= 144)] = 0.0
    fs = o.forward_star(0, 0)
    with D.Engine(p) as e:
        zones = np.arange(p.n_zones, dtype=np.int32)
        lab, pn, pl, ties, st = e.sssp_trees(0, 0, zones, cost=cost)
        assert np.all(ties > 0) and st["replayed"] == p.n_zones
        for z in zones:
            rl, rpn, rpl, _, _ = o.sssp(fs, cost, int(z))
            assert np.array_equal(lab[z], rl) and np.array_equal(pl[z], rpl) and np.array_equal(pn[z], rpn)


# ------------------------------------------------------------------------------------------------ K4
def test_vdf_known_answer(D, problems, oracle_mod):
    p = problems("corridor_101")
    o = oracle_mod.Oracle(p)
    rng = np.random.default_rng(7)
    vol = rng.uniform(0, 1.6, p.n_links) * p.cap[0] * (rng.random(p.n_links) > 0.1)
    with D.Engine(p) as e:
        out = e.vdf(0, vol)
        det = e.link_detail(0)
        speed = e.model_speed(0, 84, 13)
    exact = 0
    for l in range(0, p.n_links, 7):
        r = o.vdf_link(0, l, float(vol[l]), want_speed=True)
        for key in ("tt", "doc", "voc", "P", "vt2"):
            assert abs(out[key][l] - r[key]) <= 1e-12 * max(1.0, abs(r[key])), (key, l, out[key][l], r[key])
        exact += out["tt"][l] == r["tt"]
        for key in ("avg_queue_speed", "avg_speed_bpr", "Q_mu", "Q_gamma", "t0", "t3", "severe_congestion_P"):
            a, b = det[key][l], r[key]
            assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-9 * max(1.0, abs(b)), (key, l, a, b)
        assert np.allclose(speed[l], r["model_speed"][84:97], rtol=1e-6, atol=1e-6, equal_nan=True)
    assert exact >= 0.9 * len(range(0, p.n_links, 7))  # CUDA pow vs glibc pow: <= 2 ulp apart, mostly identical


# ------------------------------------------------------------------------------------------------ whole loop
def _compare_columns(eng_cols, ora_cols):
    assert np.array_equal(eng_cols["od_index"], ora_cols["od_index"])
    assert np.array_equal(eng_cols["node_sum"], ora_cols["node_sum"])
    assert np.array_equal(eng_cols["seq_no"], ora_cols["seq_no"])
    assert np.array_equal(eng_cols["n_links"], ora_cols["n_links"])
    assert np.array_equal(eng_cols["n_nodes"], ora_cols["n_nodes"])
    assert np.array_equal(eng_cols["links"], ora_cols["links"]), "path link sequences must be bit-exact"
    assert rel_err(eng_cols["volume"], ora_cols["volume"], floor=1e-6) <= TOL


def _sfx(tau):
    """golden keys of demand periods after the first carry a ":tau=t" suffix"""
    return "" if tau == 0 else f":tau={tau}"


@pytest.mark.parametrize("name,K,CU", [("two_corridor", 20, 0), ("corridor_101", 20, 5), ("corridor_3", 20, 20),
                                       ("corridor_3_2p", 12, 6), ("edge_cases", 10, 4), ("toll_signal", 10, 4), ("asu_qvdf", 20, 3)])
def test_assignment_matches_oracle_and_reference(D, problems, golden, oracle_mod, name, K, CU):
    p, g = problems(name), golden(name)
    o = oracle_mod.Oracle(p)
    o.capture_trees(True)
    with D.Engine(p, keep_trees=True, max_columns_per_od=K + 4) as e:
        assert e.n_tasks == o.n_tasks
        assert np.array_equal(e.task_zone, o.task_zone) and np.array_equal(e.task_at, o.task_at)
        assert np.array_equal(e.task_tau, o.task_tau)
        tree_mismatch = 0
        for k in range(K):
            r, ro = e.iteration(k), o.iteration(k)
            ref = g["iter_rows"][k]
            for key in ("sysTT", "least", "avg_tt"):
                assert abs(r[key] - ro[key]) <= TOL * max(1e-9, abs(ro[key])), (k, key, r, ro)
            assert abs(r["gap"] - ro["gap"]) <= TOL * max(1.0, abs(ro["gap"])), (k, r, ro)
            assert abs(r["gap"] - ref[3]) <= TOL * max(1.0, abs(ref[3]))
            assert r["counted_edges"] == ro["counted_edges"]
            if k in (0, 1, 2, 5, K - 1):
                lab_o, pn_o, pl_o = o._cap
                for ti in range(e.n_tasks):
                    lab, pn, pl = e.tree(ti)
                    assert rel_err(lab, lab_o[ti]) <= 1e-12
                    tree_mismatch += int(np.count_nonzero(pl != pl_o[ti]) + np.count_nonzero(pn != pn_o[ti]))
            for tau in range(p.n_periods):
                st, so = e.link_state(tau), o.link_state(tau)
                assert rel_err(st["tt"], so["tt"]) <= TOL
                assert rel_err(st["pce_volume"], so["pce_volume"], floor=1e-3) <= TOL
        assert tree_mismatch == 0, "predecessor trees must match the reference's"
        for n in range(CU):
            r, ro = e.column_update(n), o.column_update(n)
            for key in ("avg_tt", "total_gap", "rel_gap_pct", "demand"):
                assert abs(r[key] - ro[key]) <= TOL * max(1e-9, abs(ro[key])), (n, key, r, ro)
            for tau in range(p.n_periods):
                assert rel_err(e.link_state(tau)["tt"], g[f"cu{n}_tt" + _sfx(tau)]) <= TOL
        sys_tt = e.finalize(K)
        assert abs(sys_tt - float(g["final_sysTT"][0])) <= TOL * max(1.0, float(g["final_sysTT"][0]))
        for tau in range(p.n_periods):
            st = e.link_state(tau)
            for key, gk in (("tt", "final_tt"), ("pce_volume", "final_pce_volume"), ("person_volume", "final_person_volume"),
                            ("link_volume", "final_vdf_link_volume"), ("doc", "final_doc"), ("voc", "final_voc"),
                            ("P", "final_P"), ("vt2", "final_vt2")):
                assert rel_err(st[key], g[gk + _sfx(tau)], floor=1e-3) <= TOL, (key, tau)
        o.finalize(K)
        ec, oc = e.columns(), o.columns()
        _compare_columns(ec, oc)
        # and against the reference engine's own column pool
        assert np.array_equal(ec["node_sum"], g["col_node_sum"]) and np.array_equal(ec["n_links"], g["col_nlinks"])
        assert sha(ec["links"]) == str(g["col_links_sha256"])
        assert rel_err(ec["volume"], g["col_volume"], floor=1e-6) <= TOL
        if CU:
            assert rel_err(ec["cost"], g["col_cost"]) <= TOL and rel_err(ec["time"], g["col_time"]) <= TOL
        t = e.timing()
        assert t["launches"]["sssp"] >= K and t["ms"]["sssp"] > 0
        c = e.counters()
        assert c["relaxations"] > 0 and c["link_refs"] > 0 and c["path_nodes"] > 0


def test_scatter_is_order_independent_and_shards_sum_exactly(D, problems):
    """Fixed-point integer accumulation: identical bits run to run, and the shards of a 2- or 3-way origin
    partition add up to exactly the single-GPU vector (what the NCCL all-reduce computes)."""
    p = problems("corridor_101")
    def run(world, rank):
        with D.Engine(p, rank=rank, world_size=world) as e:
            if world > 1:
                e.comm_init(None)
            e.iteration(0)       # builds the first columns of this rank's origins
            e.scatter(-1)
            return e.volume_fixed(0), e.n_tasks
    full, n_full = run(1, 0)
    again, _ = run(1, 0)
    assert np.array_equal(full, again)
    assert full.sum() > 0
    for world in (2, 3):
        parts = [run(world, r) for r in range(world)]
        assert sum(n for _, n in parts) == n_full
        assert np.array_equal(sum(v for v, _ in parts), full)


def test_linearity_of_scatter(D, problems):
    """Size-independent property: scattering after an MSA decay by k/(k+1) scales every link volume by the same
    factor (up to the float rounding of each path volume the reference applies, assignment.h:87)."""
    p = problems("corridor_101")
    with D.Engine(p) as e:
        e.iteration(0)
        e.scatter(1)          # scatter, then decay volumes by 1/2
        v1 = e.link_state(0)["pce_volume"]
        e.scatter(-1)
        v2 = e.link_state(0)["pce_volume"]
    m = v1 > 1e-3
    assert m.sum() > 100 and np.allclose(v2[m], 0.5 * v1[m], rtol=1e-5)


def test_drop_in_executable_writes_reference_files(D, tmp_path):
    """network_assignment through the reference-style executable: same files, same summary rows."""
    exe = os.path.join(ROOT, "dlsim-mrm_b200", "lib", "DTALite_b200")
    src = os.path.join(FIXTURES, "two_corridor")
    for f in os.listdir(src):
        shutil.copy(os.path.join(src, f), tmp_path)
    r = subprocess.run([exe], cwd=tmp_path, stdin=subprocess.DEVNULL, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()
    assert b"done." in r.stdout
    summary = open(tmp_path / "final_summary.csv").read().splitlines()
    gold = [l.strip() for l in open(os.path.join(GOLDEN, "two_corridor_final_summary.txt")) if l[:1].isdigit()]
    rows = [l for l in summary if l[:1].isdigit()]
    assert len(rows) == 20
    for mine, ref in zip(rows, gold):
        a, b = mine.split(","), ref.split(",")
        assert a[:4] == b[:4], (mine, ref)                       # iteration, cpu, agents, avg travel time
        assert abs(float(a[4]) / 100 - float(b[4])) < 5e-6 * max(1, abs(float(b[4])))  # gap printed x100 by the committed source
    lp = open(tmp_path / "link_performance.csv").read().splitlines()
    assert lp[0].startswith("link_id,vdf_type,from_node_id,to_node_id,lanes,distance_km") and lp[0].endswith("notes")
    assert len(lp) == 5
    cols = lp[0].split(",")
    first = dict(zip(cols, lp[1].split(",")))
    assert first["link_id"] == "1" and first["vdf_type"] == "bpr" and first["time_period"] == "0700_0800"
    ra = open(tmp_path / "route_assignment.csv").read().splitlines()
    assert ra[0].startswith("first_column,path_no,o_zone_id,d_zone_id,od_pair")
    assert len(ra) == 3 and ",1->2," in ra[1]
    vols = sorted(float(dict(zip(ra[0].split(","), l.split(",")))["volume"]) for l in ra[1:])
    assert abs(sum(vols) - 7000) < 1e-3


@pytest.mark.parametrize("workload,K,CU,kernel", [("grid2k", 8, 3, "gq"), ("grid2k", 8, 3, "smq"), ("grid2k", 8, 3, "bundle"),
                                                  ("grid10k", 4, 0, "gq"), ("grid10k", 4, 0, "smq"), ("grid10k", 4, 0, "bundle"),
                                                  ("grid10k", 4, 0, "bundle16")])
def test_synthetic_grid_matches_oracle(D, oracle_mod, tmp_path, monkeypatch, workload, K, CU, kernel):
    """The bench generator at sizes the oracle finishes in seconds: same parity bars as the reference datasets."""
    _force_kernel(monkeypatch, kernel)
    from dlsim_mrm_b200 import synth
    synth.write_grid(str(tmp_path), **synth.WORKLOADS[workload])
    p = D.read_gmns(str(tmp_path), 8)
    o = oracle_mod.Oracle(p)
    o.capture_trees(True)
    with D.Engine(p, keep_trees=True, max_columns_per_od=K + 4) as e:
        for k in range(K):
            r, ro = e.iteration(k), o.iteration(k)
            for key in ("sysTT", "least"):
                assert abs(r[key] - ro[key]) <= TOL * max(1e-9, abs(ro[key])), (k, key, r, ro)
            assert abs(r["gap"] - ro["gap"]) <= TOL * max(1.0, abs(ro["gap"]))
            assert r["counted_edges"] == ro["counted_edges"]
            lab_o, pn_o, pl_o = o._cap
            for ti in range(0, e.n_tasks, 7):
                lab, pn, pl = e.tree(ti)
                assert rel_err(lab, lab_o[ti]) <= 1e-12
                assert np.array_equal(pl, pl_o[ti]) and np.array_equal(pn, pn_o[ti])
        for n in range(CU):
            r, ro = e.column_update(n), o.column_update(n)
            for key in ("avg_tt", "total_gap", "rel_gap_pct"):
                assert abs(r[key] - ro[key]) <= TOL * max(1e-9, abs(ro[key])), (n, key, r, ro)
        e.finalize(K); o.finalize(K)
        st, so = e.link_state(0), o.link_state(0)
        assert rel_err(st["tt"], so["tt"]) <= TOL and rel_err(st["pce_volume"], so["pce_volume"], floor=1e-3) <= TOL
        _compare_columns(e.columns(), o.columns())
        assert e.counters()["replayed"] == 0   # jittered lengths: no exact ties


@pytest.mark.parametrize("kernel", KERNELS)
def test_small_queues_fall_back_to_worst_case_rerun(D, problems, golden, oracle_mod, monkeypatch, kernel):
    """Trees whose frontier outgrows the regular work queues are recomputed with worst-case queues: same bits."""
    _force_kernel(monkeypatch, kernel)
    p, g = problems("corridor_101"), golden("corridor_101")
    o = oracle_mod.Oracle(p)
    cost = g["gen_cost_k2_at0"]
    fs = o.forward_star(0, 0)
    # the shared-memory-queue kernel spills and compacts, so only a far pile of a few entries makes it give up
    monkeypatch.setenv("DTL_SSSP_QUEUE_CAP", "48" if kernel == "gq" else "4")
    with D.Engine(p) as e:
        zones = np.arange(0, p.n_zones, 9, dtype=np.int32)
        lab, pn, pl, ties, st = e.sssp_trees(0, 0, zones, cost=cost)
        assert e.counters()["reruns"] > 0
        for i, z in enumerate(zones):
            rl, rpn, rpl, _, _ = o.sssp(fs, cost, int(z))
            assert np.array_equal(lab[i], rl) and np.array_equal(pl[i], rpl) and np.array_equal(pn[i], rpn)


# ------------------------------------------------------------------------------------------------ node-major trees
def _loop_vs_oracle(D, oracle_mod, p, K, CU=0):
    """whole assignment loop WITHOUT kept trees (an origin-bundle batch then leaves its trees node-major and the back-trace walks
    them there) against the oracle: columns bit-exact, link state within TOL"""
    o = oracle_mod.Oracle(p)
    with D.Engine(p, keep_trees=False, max_columns_per_od=K + 4) as e:
        for k in range(K):
            r, ro = e.iteration(k), o.iteration(k)
            for key in ("sysTT", "least"):
                assert abs(r[key] - ro[key]) <= TOL * max(1e-9, abs(ro[key])), (k, key, r, ro)
            assert abs(r["gap"] - ro["gap"]) <= TOL * max(1.0, abs(ro["gap"]))
        for n in range(CU):
            r, ro = e.column_update(n), o.column_update(n)
            assert abs(r["rel_gap_pct"] - ro["rel_gap_pct"]) <= TOL * max(1.0, abs(ro["rel_gap_pct"]))
        e.finalize(K); o.finalize(K)
        st, so = e.link_state(0), o.link_state(0)
        assert rel_err(st["tt"], so["tt"]) <= TOL and rel_err(st["pce_volume"], so["pce_volume"], floor=1e-3) <= TOL
        _compare_columns(e.columns(), o.columns())
        return e.counters(), e.sssp_kernel


@pytest.mark.parametrize("workload,kernel,tags", [("grid2k", "bundle16", 0), ("grid10k", "bundle16", 0), ("grid10k", "bundle16", 1),
                                                  ("grid2k", "bundle", 0)])
def test_node_major_trees_and_backtrace(D, oracle_mod, tmp_path, monkeypatch, workload, kernel, tags):
    _force_kernel(monkeypatch, kernel)
    if tags:
        monkeypatch.setenv("DTL_BUNDLE_TAGS", "1")   # the search variant that checks zone tags per edge (routing.h:779-786)
    from dlsim_mrm_b200 import synth
    synth.write_grid(str(tmp_path), **synth.WORKLOADS[workload])
    p = D.read_gmns(str(tmp_path), 8)
    c, used = _loop_vs_oracle(D, oracle_mod, p, 5, 2 if workload == "grid2k" else 0)
    assert "bundle" in used and c["replayed"] == 0 and c["reruns"] == 0


@pytest.mark.parametrize("name,K,CU", [("corridor_101", 6, 2), ("edge_cases", 6, 2), ("corridor_3_2p", 5, 2)])
def test_node_major_trees_on_reference_datasets(D, problems, oracle_mod, monkeypatch, name, K, CU):
    """agent-type filters, unreachable zones, two periods: several forward stars per iteration, ragged bundles"""
    _force_kernel(monkeypatch, "bundle16")
    c, used = _loop_vs_oracle(D, oracle_mod, problems(name), K, CU)
    assert "bundle" in used


def test_node_major_trees_fall_back_on_ties_and_overflow(D, oracle_mod, monkeypatch, problems):
    """a tie or an overflowed bundle found by the label -> tree pass sends the batch through the per-tree path (replay / re-run)"""
    _force_kernel(monkeypatch, "bundle16")
    c, _ = _loop_vs_oracle(D, oracle_mod, _tie_grid(), 4)
    assert c["replayed"] > 0
    monkeypatch.setenv("DTL_SSSP_QUEUE_CAP", "4")
    c, _ = _loop_vs_oracle(D, oracle_mod, problems("corridor_101"), 3)
    assert c["reruns"] > 0


# ------------------------------------------------------------------------------------------------ pipelined iterations
def _pipelined_vs_sequential(D, oracle_mod, p, K, expect_pipeline=True):
    """dtl_iterations (back-trace of iteration k beside the search of iteration k+1, two sets of tree buffers) must leave exactly what
    K calls of dtl_iteration leave: the same rows, the same Q31.32 link volumes, the same columns -- and both match the oracle."""
    o = oracle_mod.Oracle(p)
    rows_o = [o.iteration(k) for k in range(K)]
    o.finalize(K)
    with D.Engine(p, keep_trees=False, max_columns_per_od=K + 4) as e1, D.Engine(p, keep_trees=False, max_columns_per_od=K + 4) as e2:
        rows1 = [e1.iteration(k) for k in range(K)]
        rows2 = e2.iterations(0, 2) + e2.iterations(2, K - 2)      # two calls: the pipeline is entered and left twice
        for k, (r1, r2, ro) in enumerate(zip(rows1, rows2, rows_o)):
            assert r1 == r2, (k, r1, r2)                             # deterministic sums: the same bits
            for key in ("sysTT", "least"):
                assert abs(r2[key] - ro[key]) <= TOL * max(1e-9, abs(ro[key])), (k, key, r2, ro)
        assert np.array_equal(e1.volume_fixed(0), e2.volume_fixed(0))
        e1.finalize(K); e2.finalize(K)
        s1, s2, so = e1.link_state(0), e2.link_state(0), o.link_state(0)
        assert np.array_equal(s1["tt"], s2["tt"]) and np.array_equal(s1["pce_volume"], s2["pce_volume"])
        assert rel_err(s2["tt"], so["tt"]) <= TOL and rel_err(s2["pce_volume"], so["pce_volume"], floor=1e-3) <= TOL
        _compare_columns(e2.columns(), o.columns())
        c1, c2 = e1.counters(), e2.counters()
        assert c1["replayed"] == c2["replayed"] and c1["reruns"] == c2["reruns"] and c1["tasks"] == c2["tasks"]
        return c2, e2.sssp_kernel


@pytest.mark.parametrize("workload,kernel", [("grid2k", "smq"), ("grid2k", "gq"), ("grid2k", "bundle"), ("grid10k", "bundle16"), ("grid10k", "smq")])
def test_pipelined_iterations_match_sequential_and_oracle(D, oracle_mod, tmp_path, monkeypatch, workload, kernel):
    _force_kernel(monkeypatch, kernel)
    from dlsim_mrm_b200 import synth
    synth.write_grid(str(tmp_path), **synth.WORKLOADS[workload])
    p = D.read_gmns(str(tmp_path), 8)
    c, used = _pipelined_vs_sequential(D, oracle_mod, p, 6)
    assert c["replayed"] == 0 and c["reruns"] == 0


@pytest.mark.parametrize("kernel", ["bundle16", "smq"])
def test_pipelined_iterations_redo_batches_with_ties_and_overflows(D, oracle_mod, monkeypatch, tmp_path, kernel):
    """the slow paths inside the pipeline: order-dependent ties (replay) and queues that overflow (worst-case re-run)"""
    _force_kernel(monkeypatch, kernel)
    c, _ = _pipelined_vs_sequential(D, oracle_mod, _tie_grid(), 5)
    assert c["replayed"] > 0
    monkeypatch.setenv("DTL_SSSP_QUEUE_CAP", "4")
    from dlsim_mrm_b200 import synth
    synth.write_grid(str(tmp_path), **synth.WORKLOADS["grid2k"])
    c, _ = _pipelined_vs_sequential(D, oracle_mod, D.read_gmns(str(tmp_path), 8), 4)
    assert c["reruns"] > 0


@pytest.mark.parametrize("name", ["corridor_101", "corridor_3_2p"])
def test_iterations_call_on_several_forward_stars(D, problems, oracle_mod, name):
    """several agent types / periods: more than one tree batch per iteration, dtl_iterations runs them one iteration at a time"""
    _pipelined_vs_sequential(D, oracle_mod, problems(name), 5)


@pytest.mark.parametrize("kernel", ["bundle16", "smq"])
def test_pipelined_iterations_fail_cleanly(D, tmp_path, monkeypatch, kernel):
    """an error inside the pipeline (one column slot per OD cell: the second new path of a cell does not fit) comes back as an error,
    leaves no stream parked at a gate -- the engine can be closed -- and a fresh engine works afterwards"""
    _force_kernel(monkeypatch, kernel)
    from dlsim_mrm_b200 import synth
    synth.write_grid(str(tmp_path), **synth.WORKLOADS["grid2k"])
    p = D.read_gmns(str(tmp_path), 8)
    e = D.Engine(p, keep_trees=False, max_columns_per_od=1)
    with pytest.raises(D.EngineError, match="columns"):
        e.iterations(0, 8)
    e.close()
    with D.Engine(p, keep_trees=False, max_columns_per_od=12) as e2:
        rows = e2.iterations(0, 4)
        assert len(rows) == 4 and all(r["least"] > 0 for r in rows)


@pytest.mark.parametrize("workload,pipelined,kernel", [("grid2k", False, "bundle16"), ("grid10k", True, "bundle16"), ("grid10k", False, "bundle16"),
                                                       ("grid10k", True, "smq"), ("grid2k", False, "smq")])
def test_destination_bounded_search_changes_nothing_but_the_time(D, oracle_mod, tmp_path, monkeypatch, workload, pipelined, kernel):
    """The engine's default (dtl_set_bounded_search(1)): a search (origin bundles, shared-memory queues) stops when no queued node can still
    improve a destination with demand -- rows, Q31.32 link volumes and columns are the same bits as with the reference's complete trees
    (dtl_set_bounded_search(0)), and both match the oracle"""
    _force_kernel(monkeypatch, kernel)
    from dlsim_mrm_b200 import synth
    synth.write_grid(str(tmp_path), **synth.WORKLOADS[workload])
    p = D.read_gmns(str(tmp_path), 8)
    K = 6
    o = oracle_mod.Oracle(p)
    rows_o = [o.iteration(k) for k in range(K)]
    with D.Engine(p, max_columns_per_od=K + 4) as e1, D.Engine(p, max_columns_per_od=K + 4) as e2:
        assert e1.set_bounded_search(False) is True    # e1: complete trees; the setter returns the previous setting (the default: on)
        assert e2.set_bounded_search(True) is True
        rows1 = [e1.iteration(k) for k in range(K)]
        rows2 = e2.iterations(0, K) if pipelined else [e2.iteration(k) for k in range(K)]
        for k, (r1, r2, ro) in enumerate(zip(rows1, rows2, rows_o)):
            assert r1 == r2, (k, r1, r2)
            assert abs(r2["least"] - ro["least"]) <= TOL * abs(ro["least"])
        assert np.array_equal(e1.volume_fixed(0), e2.volume_fixed(0))
        c1, c2 = e1.columns(), e2.columns()
        for key in ("od_index", "node_sum", "seq_no", "links", "volume"):
            assert np.array_equal(c1[key], c2[key]), key
        _compare_columns(c2, o.columns())
        k1, k2 = e1.counters(), e2.counters()
        assert k2["replayed"] == 0 and k2["reruns"] == 0
        if workload == "grid10k":   # and it did stop early (on the 2 000-node grid the destinations reach the far side: nothing to save,
            assert k2["pops"] < k1["pops"]   # and the expansion counts of two runs of the asynchronous search differ by a few hundred)
