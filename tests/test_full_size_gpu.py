"""BASELINE.json's full size (synthetic regional grid: 100 172 nodes, 300 000 links, 2 000 zones, 400 000 OD rows) through
size-independent properties -- the oracle does not finish this size in test time:

* every sampled shortest-path tree carries its own optimality certificate: no edge of the forward star can improve a
  label (the Bellman fixed point the reference's label correcting converges to is unique), every reached node's predecessor
  link is tight in exact FP64, unreached nodes keep 1e15 / -9999, and the counted edges are the Graph500 count;
* the column pool conserves demand: after iteration k the volumes of an OD cell's columns sum to its demand (MSA weights
  1/(k+1) and k/(k+1), assignment.h:182-186, routing.h:619);
* the fixed-point link volumes are identical, bit for bit, whether one engine owns every origin or two engines own
  zone % 2 each (what two GPUs would all-reduce).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import dlsim_mrm_b200 as D
    return D


@pytest.fixture(scope="module")
def grid(D):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import ensure_workload
    return D.read_gmns(ensure_workload("grid100k"), 8)


def test_trees_carry_a_bellman_certificate_and_columns_conserve_demand(D, grid):
    p = grid
    K = 3
    with D.Engine(p, keep_trees=True, max_columns_per_od=K + 4) as e:
        for k in range(K):
            r = e.iteration(k)
        assert e.n_tasks == p.n_zones
        assert "bundle" in e.sssp_kernel              # 2 000 trees per launch: the origin-bundle kernel is the one under test here
        cost = e.gen_cost(0, 0)                      # generalized costs the last trees were computed with
        lf, lt, tag = p.link_from, p.link_to, p.link_tag
        counted = 0.0
        rng = np.random.default_rng(5)
        sample = np.unique(np.concatenate([[0, e.n_tasks - 1], rng.integers(0, e.n_tasks, 38)]))
        for ti in sample:
            z = int(e.task_zone[ti])
            lab, pn, pl = e.tree(int(ti))
            usable = (tag < 0) | (tag == z)           # only the origin zone may leave its centroid (routing.h:779-786)
            reached = lab < 1e15
            # Bellman: label[to] <= label[from] + cost for every usable edge out of a reached node, in exact FP64
            ok = usable & reached[lf]
            assert np.all(lab[lt[ok]] <= lab[lf[ok]] + cost[ok]), f"tree {ti}: an edge can still improve a label"
            # predecessors: tight, consistent, and absent exactly where the label is
            has = pl >= 0
            origin = int(p.zone_centroid[z])
            assert lab[origin] == 0.0 and not has[origin]
            assert np.array_equal(has, reached & (np.arange(p.n_nodes) != origin))
            v = np.nonzero(has)[0]
            assert np.array_equal(lt[pl[v]], v) and np.array_equal(lf[pl[v]], pn[v]) and np.all(usable[pl[v]])
            assert np.array_equal(lab[pn[v]] + cost[pl[v]], lab[v]), f"tree {ti}: a predecessor link is not tight"
            assert np.all(pl[~reached] == -9999) and np.all(pn[~reached] == -9999)
        # Graph500 edge count of the whole iteration: every tree reaches every node here, so each counts the physical
        # links plus the connectors into every centroid and the origin zone's own outgoing connector
        n_conn_in = int(np.count_nonzero(tag == -1)) - 300000
        assert r["counted_edges"] == p.n_zones * (300000 + n_conn_in + 1)
        c = e.columns()
        per_cell = np.bincount(c["od_index"], weights=c["volume"], minlength=p.n_od)
        assert np.allclose(per_cell, p.od_vol, rtol=1e-6, atol=1e-9), "column volumes do not sum to the OD demand"  # (float)(1/(k+1)), routing.h:435
        # paths are link chains from the origin centroid to the destination centroid
        first = np.concatenate([[0], np.cumsum(c["n_links"])[:-1]])
        last = first + c["n_links"] - 1
        assert np.array_equal(lf[c["links"][first]], p.zone_centroid[p.od_o[c["od_index"]]])
        assert np.array_equal(lt[c["links"][last]], p.zone_centroid[p.od_d[c["od_index"]]])


def test_sharded_engines_sum_to_the_same_fixed_point_volumes(D, grid):
    """Iterations 0 and 1 run on free-flow times on every shard (main_api.cpp:606,618), so a shard that is not all-reduced
    (shard-only mode) still builds exactly the columns it would build inside a 2-GPU run: the integer sums must agree."""
    p = grid

    def volumes(**kw):
        with D.Engine(p, max_columns_per_od=4, **kw) as e:
            if kw:
                e.comm_init(None)                     # shard-only mode: the caller sums, like the all-reduce would
            e.iteration(0)
            e.iteration(1)
            e.scatter(-1)
            return e.volume_fixed(0)

    whole = volumes()
    parts = [volumes(rank=r, world_size=2) for r in range(2)]
    assert whole.sum() > 0
    assert np.array_equal(parts[0] + parts[1], whole), "fixed-point volumes must not depend on the sharding"
