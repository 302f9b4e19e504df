"""Host-side plumbing of the origin-sharded multi-GPU run (one process per GPU).

Origin zones are partitioned into spatially contiguous blocks, one per rank (a recursive median split of a landmark
embedding of the network, equal to within one zone; ``dtl_plan_tasks`` / ``engine.cu: zone_owners``).  The reference deals
zones to its memory blocks round-robin (zone % blocks, main_api.cpp:259; ``DTL_SHARD=mod`` selects that rule); link volumes are
integer sums and therefore identical for any partition.  The only data-path collective -- the all-reduce of the fixed-point link-volume vector -- is issued
by the engine itself over NCCL.  What is here is the rendezvous around it: broadcasting the 128-byte NCCL id and
reducing timings, over whatever torch.distributed backend the launcher initialised (nccl on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import numpy as np


def broadcast_bytes(dist, payload: bytes | None, n: int, src: int = 0, device="cpu") -> bytes:
    """Broadcast ``n`` bytes from rank ``src``."""
    import torch
    if dist is None or dist.get_world_size() == 1:
        return bytes(payload)
    if dist.get_rank() == src:
        t = torch.tensor(list(payload), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(n, dtype=torch.uint8, device=device)
    dist.broadcast(t, src)
    return bytes(t.cpu().tolist())


def max_over_ranks(dist, x: float, device="cpu") -> float:
    """Timings of a multi-GPU step are the max over ranks."""
    import torch
    if dist is None or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, a: np.ndarray, device="cpu") -> np.ndarray:
    """Exact integer sum of the ranks' fixed-point volume shards (what the engine's NCCL all-reduce computes)."""
    import torch
    if dist is None or dist.get_world_size() == 1:
        return a
    t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
