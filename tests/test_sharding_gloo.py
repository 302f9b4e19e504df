"""N > 1 host path on CPU: two gloo ranks plan their shards with the native (host-only) dtl_plan_tasks, exchange the
rendezvous payload, and reduce a stand-in for the fixed-point volume vector exactly."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import FIXTURES, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dlsim_mrm_b200 as D
    from dlsim_mrm_b200 import sharding
    p = D.read_gmns(os.path.join(FIXTURES, "corridor_101"), 4)
    at, tau, zone = D.plan_tasks(p, rank, world)
    os.environ["DTL_SHARD"] = "mod"                       # the reference's round-robin rule stays selectable
    assert np.all(D.plan_tasks(p, rank, world)[2] % world == rank)
    del os.environ["DTL_SHARD"]
    payload = bytes(range(128)) if rank == 0 else None
    got = sharding.broadcast_bytes(dist, payload, 128)
    assert got == bytes(range(128))
    # every rank contributes the Q31.32 volume of "its" origins; the integer sum must not depend on the split
    q = np.zeros(p.n_links, np.int64)
    mine = np.flatnonzero(np.isin(p.od_o, zone))
    np.add.at(q, p.od_d[mine] % p.n_links, np.rint(p.od_vol[mine] * 2.0**32).astype(np.int64))
    total = sharding.sum_over_ranks(dist, q)
    t = sharding.max_over_ranks(dist, float(rank + 1))
    assert t == float(world)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), at=at, tau=tau, zone=zone, total=total)
    dist.destroy_process_group()


def test_two_rank_plan_and_exact_reduction(native, problems, tmp_path):
    import dlsim_mrm_b200 as D
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    p = problems("corridor_101")
    full = D.plan_tasks(p, 0, 1)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    key = lambda a, t, z: (a.astype(np.int64) * p.n_periods + t) * p.n_zones + z
    merged = np.sort(np.concatenate([key(q["at"], q["tau"], q["zone"]) for q in parts]))
    assert np.array_equal(merged, np.sort(key(*full)))           # disjoint cover of the single-GPU task list
    assert len(set(merged.tolist())) == merged.size
    # spatial blocks: equal to within one zone, and every rank derives the same partition on its own
    sizes = [np.unique(q["zone"]).size for q in parts]
    assert max(sizes) - min(sizes) <= 1 + (p.n_zones - np.unique(full[2]).size)
    q = np.zeros(p.n_links, np.int64)
    np.add.at(q, p.od_d % p.n_links, np.rint(p.od_vol * 2.0**32).astype(np.int64))
    assert np.array_equal(parts[0]["total"], q) and np.array_equal(parts[1]["total"], q)


def test_plan_matches_reference_block_rule(problems):
    """zones below number_of_memory_blocks always get a tree; later zones only with demand (main_api.cpp:220-264)."""
    import dlsim_mrm_b200 as D
    p = problems("corridor_101", 4)
    at, tau, zone = D.plan_tasks(p, 0, 1)
    for a in range(p.n_agent_types):
        z = zone[at == a]
        expect = [i for i in range(p.n_zones) if i < 4 or p.origin_demand[i] > 0.001]
        assert list(z) == expect
