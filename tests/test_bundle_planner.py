"""Host-side planner of the origin-bundle shortest-path kernel (dtl_plan_bundles): every origin lands in exactly one
bundle, bundles are full except at most one, and the origins of a bundle are neighbours in the network.  CPU only."""
import ctypes as C

import numpy as np


def _grid(rows, cols):
    """Directed 4-neighbour lattice as a forward star."""
    n = rows * cols
    r, c = np.divmod(np.arange(n), cols)
    eu, ev = [], []
    for dr, dc in ((0, 1), (0, -1), (1, 0), (-1, 0)):
        ok = (r + dr >= 0) & (r + dr < rows) & (c + dc >= 0) & (c + dc < cols)
        eu.append(np.arange(n)[ok]); ev.append(((r + dr) * cols + (c + dc))[ok])
    eu, ev = np.concatenate(eu), np.concatenate(ev)
    order = np.argsort(eu, kind="stable")
    row_ptr = np.zeros(n + 1, np.int32)
    np.add.at(row_ptr, eu + 1, 1)
    return np.cumsum(row_ptr).astype(np.int32), ev[order].astype(np.int32), n


def _plan(native, row_ptr, head, n_nodes, origins, B):
    native.dtl_plan_bundles.restype = C.c_int
    native.dtl_plan_bundles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    cap = (len(origins) // B + 2) * B
    out = np.full(cap, -7, np.int32)
    nb = native.dtl_plan_bundles(row_ptr.ctypes.data, head.ctypes.data, n_nodes, origins.ctypes.data, len(origins), B,
                                 out.ctypes.data, cap)
    assert nb > 0
    return out[: nb * B].reshape(nb, B)


def test_bundles_partition_the_origins_and_are_compact(native):
    rows, cols = 60, 70
    row_ptr, head, n = _grid(rows, cols)
    rng = np.random.default_rng(3)
    origins = rng.permutation(n)[:403].astype(np.int32)         # ragged: 403 = 25 * 16 + 3
    for B in (4, 8, 16):
        b = _plan(native, row_ptr, head, n, origins, B)
        used = b[b >= 0]
        assert np.array_equal(np.sort(used), np.arange(len(origins))), "every origin exactly once"
        full = (b >= 0).sum(axis=1)
        assert b.shape[0] == -(-len(origins) // B) and np.count_nonzero(full < B) <= 1
        # compactness: the mean pairwise lattice distance inside a bundle is far below that of bundles of consecutive origins
        r, c = np.divmod(origins, cols)

        def spread(groups):
            tot, cnt = 0.0, 0
            for gidx in groups:
                gidx = gidx[gidx >= 0]
                d = np.abs(r[gidx][:, None] - r[gidx][None, :]) + np.abs(c[gidx][:, None] - c[gidx][None, :])
                tot += d.sum(); cnt += d.size
            return tot / cnt
        consecutive = [np.arange(i, min(i + B, len(origins))) for i in range(0, len(origins), B)]
        assert spread(list(b)) < 0.35 * spread(consecutive), (B, spread(list(b)), spread(consecutive))


def test_planner_rejects_bad_arguments(native):
    row_ptr, head, n = _grid(4, 4)
    origins = np.array([0, 5, 99], np.int32)
    native.dtl_plan_bundles.restype = C.c_int
    native.dtl_plan_bundles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    out = np.zeros(32, np.int32)
    assert native.dtl_plan_bundles(row_ptr.ctypes.data, head.ctypes.data, n, origins.ctypes.data, 3, 16, out.ctypes.data, 32) < 0
    assert native.dtl_plan_bundles(row_ptr.ctypes.data, head.ctypes.data, n, origins.ctypes.data, 2, 5, out.ctypes.data, 32) < 0
