"""The multi-GPU path over NCCL on real devices (needs >= 2 GPUs; skipped otherwise): two ranks own spatial halves of the origin
zones of BASELINE.json config 4 (4_101_corridor, three agent types), all-reduce the fixed-point link volumes once per iteration
and must hold, bit for bit, the vector a single GPU computes -- and a failure on one rank must end the iteration on BOTH
ranks with an error instead of leaving the peer blocked in its next collective (raw NCCL has no watchdog)."""
import hashlib
import os
import sys
import time

import numpy as np
import pytest

from conftest import FIXTURES, ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _exchange_id(rank, path, D):
    """rank 0 makes the NCCL id; the others read it from a file (what bench.py does with a torch.distributed broadcast)"""
    if rank == 0:
        uid = D.Engine.comm_unique_id()
        with open(path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(path + ".tmp", path)
        return uid
    for _ in range(600):
        if os.path.exists(path):
            return open(path, "rb").read()
        time.sleep(0.05)
    raise RuntimeError("no NCCL id from rank 0")


def _worker(rank, world, out_dir, mode):
    sys.path.insert(0, ROOT)
    os.environ["DTL_NCCL_TIMEOUT_S"] = "30"
    import dlsim_mrm_b200 as D
    p = D.read_gmns(os.path.join(FIXTURES, "corridor_101"), 4)
    uid = _exchange_id(rank, os.path.join(out_dir, f"id_{mode}"), D)
    cap = 1 if (mode == "fail" and rank == 1) else 12   # one column slot per OD cell: rank 1 runs out of slots in the loop
    res = {"error": "", "failed_at": -1}
    with D.Engine(p, device=rank, rank=rank, world_size=world, max_columns_per_od=cap) as e:
        e.comm_init(uid)
        try:
            for k in range(6):
                res["failed_at"] = k
                e.iteration(k)
            res["failed_at"] = -1
            e.scatter(-1)
            res["volume"] = e.volume_fixed(0)
            res["n_tasks"] = e.n_tasks
        except D.EngineError as ex:
            res["error"] = str(ex)
    np.savez(os.path.join(out_dir, f"{mode}_rank{rank}.npz"), **res)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_all_reduce_the_single_gpu_volumes(tmp_path, world):
    """BASELINE.json config 4 origin-sharded over 2, 4 and 8 GPUs (99 zones x 3 agent types: 12-13 zones per GPU at 8)"""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import dlsim_mrm_b200 as D
    mp.spawn(_worker, args=(world, str(tmp_path), "ok"), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"ok_rank{r}.npz") for r in range(world)]
    assert all(str(q["error"]) == "" for q in parts), [str(q["error"]) for q in parts]
    p = D.read_gmns(os.path.join(FIXTURES, "corridor_101"), 4)
    with D.Engine(p, device=0, max_columns_per_od=12) as e:
        for k in range(6):
            e.iteration(k)
        e.scatter(-1)
        whole, n_whole = e.volume_fixed(0), e.n_tasks
    assert sum(int(q["n_tasks"]) for q in parts) == n_whole
    digest = hashlib.sha256(whole.tobytes()).hexdigest()
    for q in parts:   # after the all-reduce every rank holds the complete vector
        assert hashlib.sha256(q["volume"].tobytes()).hexdigest() == digest


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_a_failing_rank_stops_every_rank(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    t0 = time.time()
    mp.spawn(_worker, args=(world, str(tmp_path), "fail"), nprocs=world, join=True)
    assert time.time() - t0 < 120, "a rank was left waiting in a collective"
    parts = [np.load(tmp_path / f"fail_rank{r}.npz") for r in range(world)]
    assert all(str(q["error"]) != "" for q in parts), "both ranks must report the failure"
    assert int(parts[0]["failed_at"]) == int(parts[1]["failed_at"]) >= 0, "and stop in the same iteration"
    assert "max_columns_per_od" in str(parts[1]["error"])


def _worker_grid(rank, world, out_dir, grid_dir, mode):
    """one agent type, one batch of trees per iteration: dtl_iterations runs as a pipeline (collectives on the side stream)"""
    sys.path.insert(0, ROOT)
    os.environ["DTL_NCCL_TIMEOUT_S"] = "30"
    import dlsim_mrm_b200 as D
    p = D.read_gmns(grid_dir, 8)
    uid = _exchange_id(rank, os.path.join(out_dir, f"id_{mode}"), D)
    cap = 1 if (mode == "pfail" and rank == 1) else 12
    res = {"error": ""}
    with D.Engine(p, device=rank, rank=rank, world_size=world, max_columns_per_od=cap) as e:
        e.comm_init(uid)
        try:
            rows = e.iterations(0, 6)
            res["least"] = np.array([r["least"] for r in rows])
            res["volume"] = e.volume_fixed(0)
        except D.EngineError as ex:
            res["error"] = str(ex)
    np.savez(os.path.join(out_dir, f"{mode}_rank{rank}.npz"), **res)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_pipelined_iterations_over_nccl(tmp_path):
    """dtl_iterations on two ranks: the same rows and the same Q31.32 link volumes as one GPU running dtl_iteration six times"""
    import torch.multiprocessing as mp
    import dlsim_mrm_b200 as D
    from dlsim_mrm_b200 import synth
    grid = str(tmp_path / "grid")
    os.makedirs(grid)
    synth.write_grid(grid, **synth.WORKLOADS["grid2k"])
    mp.spawn(_worker_grid, args=(2, str(tmp_path), grid, "pipe"), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"pipe_rank{r}.npz") for r in range(2)]
    assert all(str(q["error"]) == "" for q in parts), [str(q["error"]) for q in parts]
    p = D.read_gmns(grid, 8)
    with D.Engine(p, device=0, max_columns_per_od=12) as e:
        rows = [e.iteration(k) for k in range(6)]
        whole = e.volume_fixed(0)
    for q in parts:
        assert np.array_equal(q["volume"], whole)
        assert np.allclose(q["least"], [r["least"] for r in rows], rtol=1e-12)   # an all-reduced FP64 sum of two partial sums


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_a_failing_rank_stops_every_rank_inside_the_pipeline(tmp_path):
    import torch.multiprocessing as mp
    from dlsim_mrm_b200 import synth
    grid = str(tmp_path / "grid")
    os.makedirs(grid)
    synth.write_grid(grid, **synth.WORKLOADS["grid2k"])
    t0 = time.time()
    mp.spawn(_worker_grid, args=(2, str(tmp_path), grid, "pfail"), nprocs=2, join=True)
    assert time.time() - t0 < 150, "a rank was left waiting in a collective"
    parts = [np.load(tmp_path / f"pfail_rank{r}.npz") for r in range(2)]
    assert all(str(q["error"]) != "" for q in parts), "both ranks must report the failure"
