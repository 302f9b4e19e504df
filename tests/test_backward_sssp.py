"""the backward per-destination search of the reference's real-time-information layer
(NetworkForSP::optimal_backward_label_correcting_from_destination, routing.h:1009-1273).
  * CPU: the oracle restatement is pinned bit for bit against trees dumped from the reference itself
    (tests/golden/backward_ref.npz, made by `make_golden.py backward`).
  * GPU: dtl_backward_trees (the K1 kernels on the in-link star) against the oracle and the same dumps, every K1 kernel."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN

CASES = ["two_corridor", "edge_cases", "corridor_101", "toll_signal"]


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def bw():
    return np.load(os.path.join(GOLDEN, "backward_ref.npz"))


def _zones(bw, name):
    """(zone, sha) of every dumped tree of agent type 0 (the star does not depend on the agent type without an RT mask)"""
    out = []
    for key, digest in zip(bw[f"{name}:keys"], bw[f"{name}:sha256"]):
        at, tau, z = (int(x.split("=")[1]) for x in str(key).split(":"))
        out.append((at, z, str(digest)))
    return out


@pytest.mark.parametrize("name", CASES)
def test_oracle_backward_trees_match_the_reference(problems, oracle_mod, bw, name):
    p = problems(name)
    o = oracle_mod.Oracle(p)
    cost = bw[f"{name}:cost"]
    n = 0
    for at, z, digest in _zones(bw, name):
        lab, pn, pl, pops = o.backward_sssp(cost, int(p.zone_centroid[z]))
        assert _sha(lab) + _sha(pl) + _sha(pn) == digest, f"{name}: backward tree of zone {z} (agent type {at}) differs"
        full = f"{name}:label:at={at}:tau=0:z={z}"
        if full in bw.files:
            assert np.array_equal(lab, bw[full]) and np.array_equal(pl, bw[f"{name}:pred_link:at={at}:tau=0:z={z}"])
            assert np.array_equal(pn, bw[f"{name}:pred_node:at={at}:tau=0:z={z}"])
            assert np.count_nonzero(lab < 1e15) > 1   # the destination is reached from somewhere
        n += 1
    assert n >= 2


def test_oracle_backward_mask_and_tags(problems, oracle_mod, bw):
    """links with an outgoing-connector tag are never used (routing.h:1104), masked links neither (routing.h:1119)"""
    p = problems("edge_cases")
    o = oracle_mod.Oracle(p)
    cost = bw["edge_cases:cost"]
    dest = int(p.zone_centroid[0])
    lab, pn, pl, _ = o.backward_sssp(cost, dest)
    used = pl[pl >= 0]
    assert np.all(p.link_tag[used] < 0)
    assert np.all(p.link_to[used] == pn[pl >= 0])
    # closing the destination's first in-link changes (or removes) the paths through it
    first_in = int(np.flatnonzero(p.link_to == dest)[0])
    mask = np.ones(p.n_links, np.uint8); mask[first_in] = 0
    lab2, pn2, pl2, _ = o.backward_sssp(cost, dest, rt_allowed=mask)
    assert not np.any(pl2 == first_in) and np.all(lab2 >= lab)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["gq", "smq", "bundle", "bundle16"])
@pytest.mark.parametrize("name", CASES)
def test_gpu_backward_trees_bit_exact(problems, oracle_mod, bw, monkeypatch, name, kernel):
    import dlsim_mrm_b200 as D
    monkeypatch.setenv("DTL_SSSP_KERNEL", "bundle" if kernel.startswith("bundle") else kernel)
    if kernel == "bundle16":
        monkeypatch.setenv("DTL_BUNDLE_B", "16")
    p = problems(name)
    o = oracle_mod.Oracle(p)
    cost = bw[f"{name}:cost"]
    dest = np.array([int(p.zone_centroid[z]) for z in range(p.n_zones)], np.int32)
    with D.Engine(p) as e:
        lab, pn, pl, ties, st = e.backward_trees(0, dest, cost=cost)
        for z in range(p.n_zones):
            rl, rpn, rpl, _ = o.backward_sssp(cost, int(dest[z]))
            assert np.array_equal(lab[z].view(np.int64), rl.view(np.int64)), f"labels differ, zone {z}"
            assert np.array_equal(pl[z], rpl) and np.array_equal(pn[z], rpn), f"predecessors differ, zone {z}"
        for at, z, digest in _zones(bw, name):   # and against the reference's own dumps
            assert _sha(lab[z]) + _sha(pl[z]) + _sha(pn[z]) == digest
        # an RT_allowed_use mask (routing.h:1119): built for the call, checked against the oracle
        mask = np.ones(p.n_links, np.uint8); mask[::7] = 0
        lab, pn, pl, ties, st = e.backward_trees(0, dest[:5], cost=cost, rt_allowed=mask)
        for i in range(min(5, p.n_zones)):
            rl, rpn, rpl, _ = o.backward_sssp(cost, int(dest[i]), rt_allowed=mask)
            assert np.array_equal(lab[i].view(np.int64), rl.view(np.int64)) and np.array_equal(pl[i], rpl) and np.array_equal(pn[i], rpn)


@pytest.mark.gpu
def test_gpu_backward_trees_on_current_travel_times(problems, oracle_mod):
    """cost = NULL: the engine's travel times after a few assignment iterations"""
    import dlsim_mrm_b200 as D
    p = problems("corridor_101")
    o = oracle_mod.Oracle(p)
    with D.Engine(p) as e:
        for k in range(3):
            e.iteration(k)
        e.finalize(3)
        tt = e.link_state(0)["tt"]
        dest = np.array([int(p.zone_centroid[z]) for z in (0, 17, 98)], np.int32)
        lab, pn, pl, ties, st = e.backward_trees(0, dest)
        for i in range(3):
            rl, rpn, rpl, _ = o.backward_sssp(tt, int(dest[i]))
            assert np.array_equal(lab[i].view(np.int64), rl.view(np.int64)) and np.array_equal(pl[i], rpl) and np.array_equal(pn[i], rpn)
