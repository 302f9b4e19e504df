"""ctypes mirror of the engine C ABI (``include/dtalite_b200.h``).

Thin by design: every method is one call into ``lib/libdtalite_b200.so``; there is no Python or CPU
compute path, and loading fails loudly when the native library is missing.  Method names follow the
reference functions they replace (``iteration`` = loop body of ``network_assignment``,
``main_api.cpp:589-728``; ``column_update`` = ``g_update_gradient_cost_and_assigned_flow_in_column_pool``,
``assignment.h:565``; ``sssp_trees`` = ``NetworkForSP::optimal_label_correcting``, ``routing.h:633``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .problem import ARRAY_SPECS, Problem

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libdtalite_b200.so")


class EngineError(RuntimeError):
    pass


class _Problem(C.Structure):
    _fields_ = (
        [(n, C.c_int) for n in ("n_nodes", "n_links", "n_zones", "n_agent_types", "n_periods", "memory_blocks")]
        + [(n, C.c_void_p) for n in ("link_from", "link_to", "link_tag", "node_zone", "zone_centroid")]
        + [(n, C.c_void_p) for n in ("fftt", "cap", "alpha", "beta", "vf", "vc", "nlanes", "cap_lane", "qdf",
                                     "preload", "penalty", "Q_alpha", "Q_beta", "Q_cd", "Q_n", "Q_cp", "Q_s", "t2",
                                     "L_h", "start_h", "end_h", "plf", "cycle_length", "red_time", "sat_flow",
                                     "vdf_type", "pce", "occ", "toll", "lr_price", "allowed", "vot")]
        + [("n_od", C.c_int)]
        + [(n, C.c_void_p) for n in ("od_o", "od_d", "od_at", "od_tau", "od_vol", "origin_demand")]
        + [("total_demand_volume", C.c_float)]
    )


class _Options(C.Structure):
    _fields_ = [("device", C.c_int), ("max_columns_per_od", C.c_int), ("rank", C.c_int), ("world_size", C.c_int),
                ("sssp_delta_factor", C.c_double), ("sssp_block_threads", C.c_int), ("sssp_ctas_per_sm", C.c_int),
                ("tree_batch_bytes", C.c_size_t), ("column_arena_bytes", C.c_size_t), ("keep_trees", C.c_int)]


_lib = None


def load_library(path: str | None = None):
    """Load the native library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("DTL_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise EngineError(f"native library {p} is missing: run `python -m dlsim_mrm_b200.build` "
                          "(there is no CPU or Python fallback)")
    L = C.CDLL(p)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    L.dtl_last_error.restype = C.c_char_p
    L.dtl_version.restype = C.c_char_p
    L.dtl_default_options.argtypes = [vp]
    L.dtl_create.restype = vp
    L.dtl_create.argtypes = [vp, vp]
    L.dtl_destroy.argtypes = [vp]
    L.dtl_comm_get_unique_id.argtypes = [vp]
    L.dtl_comm_init.argtypes = [vp, vp]
    L.dtl_plan_tasks.argtypes = [vp, ci, ci, vp, vp, vp, ci]
    L.dtl_num_tasks.argtypes = [vp]
    L.dtl_get_tasks.argtypes = [vp] * 4
    L.dtl_iteration.argtypes = [vp, ci, vp]
    L.dtl_column_update.argtypes = [vp, ci, vp]
    L.dtl_finalize.argtypes = [vp, ci, vp]
    L.dtl_run_assignment.argtypes = [vp, ci, ci, vp, vp]
    L.dtl_get_link_state.argtypes = [vp, ci] + [vp] * 8
    L.dtl_get_volume_fixed.argtypes = [vp, ci, vp]
    L.dtl_get_person_volume_at.argtypes = [vp, ci, ci, vp]
    L.dtl_get_link_detail.argtypes = [vp, ci] + [vp] * 7
    L.dtl_get_model_speed.argtypes = [vp, ci, ci, ci, vp]
    L.dtl_get_gen_cost.argtypes = [vp, ci, ci, vp]
    L.dtl_num_columns.restype = C.c_longlong
    L.dtl_num_columns.argtypes = [vp]
    L.dtl_num_column_links.restype = C.c_longlong
    L.dtl_num_column_links.argtypes = [vp]
    L.dtl_get_columns.argtypes = [vp] * 10
    L.dtl_sssp_trees.argtypes = [vp, ci, ci, vp, vp, ci, vp, vp, vp, vp, vp]
    L.dtl_backward_trees.argtypes = [vp, ci, vp, vp, vp, ci, vp, vp, vp, vp, vp]
    L.dtl_vdf.argtypes = [vp, ci] + [vp] * 6
    L.dtl_odme_configure.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp]
    L.dtl_odme_scatter.argtypes = [vp, ci, vp]
    L.dtl_odme_update.argtypes = [vp]
    L.dtl_odme_run.argtypes = [vp, ci, vp, vp]
    L.dtl_odme_get.argtypes = [vp, ci, vp, vp, vp]
    L.dtl_sa_configure.argtypes = [vp, vp, ci, vp, vp, vp]
    L.dtl_sa_begin.argtypes = [vp, vp]
    L.dtl_sa_iteration.argtypes = [vp, ci, vp]
    L.dtl_sa_column_update.argtypes = [vp, ci, vp]
    L.dtl_sa_get_columns.argtypes = [vp, vp, vp, vp, vp]
    L.dtl_gmns_scenario.argtypes = [vp, vp, vp, vp, vp, ci]
    L.dtl_gmns_measurements.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci]
    L.dtl_scatter.argtypes = [vp, ci]
    L.dtl_get_tree.argtypes = [vp, ci, vp, vp, vp]
    L.dtl_get_timing.argtypes = [vp, vp, vp, ci]
    L.dtl_get_counters.argtypes = [vp, vp, ci]
    L.dtl_device_bytes.restype = C.c_size_t
    L.dtl_device_bytes.argtypes = [vp]
    L.dtl_gmns_read.restype = vp
    L.dtl_gmns_read.argtypes = [C.c_char_p, ci]
    L.dtl_gmns_problem.restype = C.POINTER(_Problem)
    L.dtl_gmns_problem.argtypes = [vp]
    L.dtl_gmns_free.argtypes = [vp]
    L.dtl_gmns_last_error.restype = C.c_char_p
    L.dtl_gmns_ids.argtypes = [vp, vp, vp]
    L.dtl_gmns_assignment_settings.argtypes = [C.c_char_p, vp]
    L.dtl_gmns_write_outputs.argtypes = [vp, vp, C.c_char_p]
    L.network_assignment.restype = cd
    L.network_assignment.argtypes = [ci] * 7
    L.dtalite_main.restype = ci
    if path is None:
        _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c_problem(p: Problem, keep: dict) -> _Problem:
    cp = _Problem()
    for n in ("n_nodes", "n_links", "n_zones", "n_agent_types", "n_periods", "memory_blocks"):
        setattr(cp, n, int(getattr(p, n)))
    for name in ARRAY_SPECS:
        a = np.ascontiguousarray(p.arrays[name], dtype=ARRAY_SPECS[name][0])
        keep[name] = a
        setattr(cp, name, _ptr(a))
    cp.n_od = p.n_od
    cp.total_demand_volume = float(p.total_demand_volume)
    return cp


def plan_tasks(problem: Problem, rank: int, world_size: int):
    """Host-only task partition of one rank: arrays (agent type, period, zone)."""
    L = load_library()
    keep = {}
    cp = _c_problem(problem, keep)
    n = L.dtl_plan_tasks(C.byref(cp), rank, world_size, None, None, None, 0)
    if n < 0:
        raise EngineError(L.dtl_last_error().decode())
    at = np.zeros(max(n, 1), np.int32); tau = np.zeros_like(at); zone = np.zeros_like(at)
    L.dtl_plan_tasks(C.byref(cp), rank, world_size, _ptr(at), _ptr(tau), _ptr(zone), n)
    return at[:n], tau[:n], zone[:n]


def read_gmns(directory: str, memory_blocks: int = 4) -> Problem:
    """GMNS files -> Problem through the native reader (``dtl_gmns_read``; reference ``g_read_input_data``,
    ``input.h:2388``, ``g_ReadDemandFileBasedOnDemandFileList``, ``input.h:392``)."""
    L = load_library()
    h = L.dtl_gmns_read(os.fsencode(directory), int(memory_blocks))
    if not h:
        raise EngineError(L.dtl_gmns_last_error().decode())
    try:
        cp = L.dtl_gmns_problem(h).contents
        N, Lk, Z, A, T = cp.n_nodes, cp.n_links, cp.n_zones, cp.n_agent_types, cp.n_periods
        q = Problem(N, Lk, Z, A, T, cp.memory_blocks, float(cp.total_demand_volume))
        n_od = cp.n_od
        dims = {"N": N, "L": Lk, "Z": Z, "A": A, "TL": T * Lk, "TAL": T * A * Lk, "OD": n_od}
        for name, (dt, code) in ARRAY_SPECS.items():
            cnt = dims[code]
            addr = getattr(cp, name)
            if cnt == 0 or not addr:
                q.arrays[name] = np.zeros(0, dtype=dt)
            else:
                buf = (C.c_char * (cnt * np.dtype(dt).itemsize)).from_address(addr)
                q.arrays[name] = np.frombuffer(buf, dtype=dt, count=cnt).copy()
        node_id = np.zeros(N, np.int32)
        zone_id = np.zeros(Z, np.int32)
        L.dtl_gmns_ids(h, _ptr(node_id), _ptr(zone_id))
        q.node_id, q.zone_id = node_id, zone_id
        # sensitivity-analysis inputs: [agent_type] real_time_info and the "sa" rows of supply_side_scenario.csv
        rt = np.zeros(max(A, 1), np.int32)
        n_sa = L.dtl_gmns_scenario(h, _ptr(rt), None, None, None, 0)
        sa_tau, sa_link, sa_lanes = np.zeros(max(n_sa, 1), np.int32), np.zeros(max(n_sa, 1), np.int32), np.zeros(max(n_sa, 1))
        L.dtl_gmns_scenario(h, None, _ptr(sa_tau), _ptr(sa_link), _ptr(sa_lanes), n_sa)
        q.rt_info = rt[:A].copy()
        q.sa_changes = [(int(sa_tau[i]), int(sa_link[i]), float(sa_lanes[i])) for i in range(n_sa)]
        L.dtl_gmns_dms.restype = C.c_int
        L.dtl_gmns_dms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        n_dms = L.dtl_gmns_dms(h, None, None, None, 0)
        d_tau = np.zeros(max(1, n_dms), np.int32); d_link = np.zeros(max(1, n_dms), np.int32); d_zone = np.zeros(max(1, n_dms), np.int32)
        L.dtl_gmns_dms(h, _ptr(d_tau), _ptr(d_link), _ptr(d_zone), n_dms)
        q.dms_links = [(int(d_tau[i]), int(d_link[i]), int(d_zone[i])) for i in range(n_dms)]
        # ODME inputs: agent-type PCE / occupancy (float) and the link rows of measurement.csv
        pce, occ = np.zeros(max(A, 1), np.float32), np.zeros(max(A, 1), np.float32)
        n_ms = L.dtl_gmns_measurements(h, _ptr(pce), _ptr(occ), None, None, None, None, 0)
        m_tau, m_link, m_ub = (np.zeros(max(n_ms, 1), np.int32) for _ in range(3))
        m_cnt = np.zeros(max(n_ms, 1), np.float32)
        L.dtl_gmns_measurements(h, None, None, _ptr(m_tau), _ptr(m_link), _ptr(m_cnt), _ptr(m_ub), n_ms)
        q.at_pce, q.at_occ = pce[:A].copy(), occ[:A].copy()
        q.sensors = [(int(m_tau[i]), int(m_link[i]), float(m_cnt[i]), int(m_ub[i])) for i in range(n_ms)]
        return q.finalize()
    finally:
        L.dtl_gmns_free(h)


class Engine:
    """One device-resident assignment problem on one GPU (one rank of a possibly sharded run)."""

    TIMING_CLASSES = ("vdf", "scatter", "sssp", "tie_replay", "backtrace", "column_pool", "all_reduce", "iteration")
    COUNTERS = ("counted_edges", "relaxations", "pops", "rounds", "tasks", "replayed", "link_refs", "path_nodes", "reruns",
                "tie_candidates")

    def __init__(self, problem: Problem, device: int = -1, rank: int = 0, world_size: int = 1, keep_trees: bool = False,
                 max_columns_per_od: int = 0, delta_factor: float = 0.0, block_threads: int = 0, ctas_per_sm: int = 0,
                 tree_batch_bytes: int = 0, column_arena_bytes: int = 0):
        self.lib = load_library()
        self.problem = problem
        self._keep = {}
        cp = _c_problem(problem, self._keep)
        opt = _Options()
        self.lib.dtl_default_options(C.byref(opt))
        opt.device = device
        opt.rank, opt.world_size = rank, world_size
        opt.keep_trees = int(keep_trees)
        if max_columns_per_od:
            opt.max_columns_per_od = max_columns_per_od
        opt.sssp_delta_factor = delta_factor
        opt.sssp_block_threads = block_threads
        opt.sssp_ctas_per_sm = ctas_per_sm
        opt.tree_batch_bytes = tree_batch_bytes
        opt.column_arena_bytes = column_arena_bytes
        self.h = self.lib.dtl_create(C.byref(cp), C.byref(opt))
        if not self.h:
            raise EngineError(self.lib.dtl_last_error().decode())
        self.n_tasks = self.lib.dtl_num_tasks(self.h)
        at = np.zeros(max(self.n_tasks, 1), np.int32); tau = np.zeros_like(at); zone = np.zeros_like(at)
        self.lib.dtl_get_tasks(self.h, _ptr(at), _ptr(tau), _ptr(zone))
        self.task_at, self.task_tau, self.task_zone = at[:self.n_tasks], tau[:self.n_tasks], zone[:self.n_tasks]

    # -- lifetime
    def close(self):
        if getattr(self, "h", None):
            self.lib.dtl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(self.lib.dtl_last_error().decode())

    # -- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        L = load_library()
        buf = (C.c_ubyte * 128)()
        if L.dtl_comm_get_unique_id(buf) != 0:
            raise EngineError(L.dtl_last_error().decode())
        return bytes(buf)

    def comm_init(self, unique_id: bytes | None):
        """``None`` = shard-only mode (no collective; the caller sums the ranks' fixed-point volumes)."""
        if unique_id is None:
            self._ck(self.lib.dtl_comm_init(self.h, None))
            return
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._ck(self.lib.dtl_comm_init(self.h, buf))

    # -- assignment loop
    def iteration(self, k: int) -> dict:
        row = np.zeros(5)
        self._ck(self.lib.dtl_iteration(self.h, int(k), _ptr(row)))
        return dict(sysTT=row[0], least=row[1], gap=row[2], avg_tt=row[3], counted_edges=row[4])

    def iterations(self, k0: int, n: int) -> list:
        """Iterations k0 .. k0+n-1 in one call (pipelined on the device: dtl_iterations); the rows of iteration()."""
        rows = np.zeros((max(n, 1), 5))
        self.lib.dtl_iterations.restype = C.c_int
        self.lib.dtl_iterations.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self._ck(self.lib.dtl_iterations(self.h, int(k0), int(n), _ptr(rows)))
        return [dict(sysTT=r[0], least=r[1], gap=r[2], avg_tt=r[3], counted_edges=r[4]) for r in rows[:n]]

    def column_update(self, n: int) -> dict:
        row = np.zeros(4)
        self._ck(self.lib.dtl_column_update(self.h, int(n), _ptr(row)))
        return dict(avg_tt=row[0], total_gap=row[1], rel_gap_pct=row[2], demand=row[3])

    def finalize(self, K: int) -> float:
        v = C.c_double(0)
        self._ck(self.lib.dtl_finalize(self.h, int(K), C.byref(v)))
        return v.value

    def run_assignment(self, K: int, CU: int = 0):
        it = np.zeros((max(K, 1), 5)); cu = np.zeros((max(CU, 1), 4))
        self._ck(self.lib.dtl_run_assignment(self.h, int(K), int(CU), _ptr(it), _ptr(cu)))
        return it[:K], cu[:CU]

    # -- state
    def link_state(self, tau: int = 0, fields=None) -> dict:
        """per-link state of a period; ``fields`` selects the arrays to read back (default: all eight)"""
        L = self.problem.n_links
        names = ("tt", "pce_volume", "person_volume", "link_volume", "doc", "voc", "P", "vt2")
        arrs = [np.empty(L) if fields is None or n in fields else None for n in names]
        self._ck(self.lib.dtl_get_link_state(self.h, tau, *[_ptr(a) for a in arrs]))
        return {n: a for n, a in zip(names, arrs) if a is not None}

    def volume_fixed(self, tau: int = 0) -> np.ndarray:
        q = np.zeros(self.problem.n_links, np.int64)
        self._ck(self.lib.dtl_get_volume_fixed(self.h, tau, _ptr(q)))
        return q

    def person_volume_at(self, tau: int, at: int) -> np.ndarray:
        v = np.zeros(self.problem.n_links)
        self._ck(self.lib.dtl_get_person_volume_at(self.h, tau, at, _ptr(v)))
        return v

    def link_detail(self, tau: int = 0) -> dict:
        L = self.problem.n_links
        names = ("avg_queue_speed", "avg_speed_bpr", "Q_mu", "Q_gamma", "t0", "t3", "severe_congestion_P")
        arrs = [np.zeros(L) for _ in names]
        self._ck(self.lib.dtl_get_link_detail(self.h, tau, *[_ptr(a) for a in arrs]))
        return dict(zip(names, arrs))

    def model_speed(self, tau: int, first_slot: int, n_slots: int) -> np.ndarray:
        sp = np.zeros((self.problem.n_links, n_slots), np.float32)
        self._ck(self.lib.dtl_get_model_speed(self.h, tau, first_slot, n_slots, _ptr(sp)))
        return sp

    def gen_cost(self, at: int, tau: int = 0) -> np.ndarray:
        c = np.zeros(self.problem.n_links)
        self._ck(self.lib.dtl_get_gen_cost(self.h, at, tau, _ptr(c)))
        return c

    def columns(self) -> dict:
        n = self.lib.dtl_num_columns(self.h)
        nl = self.lib.dtl_num_column_links(self.h)
        if n < 0 or nl < 0:
            raise EngineError(self.lib.dtl_last_error().decode())
        i32 = lambda m: np.zeros(max(int(m), 1), np.int32)
        f64 = lambda m: np.zeros(max(int(m), 1), np.float64)
        od, ns, sq, nlk, nnd, vol, cost, tm, links = i32(n), i32(n), i32(n), i32(n), i32(n), f64(n), f64(n), f64(n), i32(nl)
        self._ck(self.lib.dtl_get_columns(self.h, *[_ptr(a) for a in (od, ns, sq, nlk, nnd, vol, cost, tm, links)]))
        return dict(od_index=od[:n], node_sum=ns[:n], seq_no=sq[:n], n_links=nlk[:n], n_nodes=nnd[:n], volume=vol[:n],
                    cost=cost[:n], time=tm[:n], links=links[:nl])

    # -- kernel-level entry points
    def sssp_trees(self, at: int, tau: int, zones, cost=None, want_trees: bool = True):
        zones = np.ascontiguousarray(zones, np.int32)
        n, N = zones.size, self.problem.n_nodes
        cost_a = None if cost is None else np.ascontiguousarray(cost, np.float64)
        lab = np.zeros((n, N)) if want_trees else None
        pn = np.zeros((n, N), np.int32) if want_trees else None
        pl = np.zeros((n, N), np.int32) if want_trees else None
        ties = np.zeros(max(n, 1), np.int32)
        stats = np.zeros(6)
        self._ck(self.lib.dtl_sssp_trees(self.h, at, tau, _ptr(cost_a), _ptr(zones), n, _ptr(lab), _ptr(pn), _ptr(pl),
                                         _ptr(ties), _ptr(stats)))
        st = dict(ms=stats[0], counted_edges=stats[1], relaxations=stats[2], pops=stats[3], rounds=stats[4],
                  replayed=stats[5])
        return lab, pn, pl, ties[:n], st

    def backward_trees(self, tau: int, dest_nodes, cost=None, rt_allowed=None):
        """Backward search from destination nodes (``optimal_backward_label_correcting_from_destination``, ``routing.h:1009``)."""
        dest = np.ascontiguousarray(dest_nodes, np.int32)
        n, N = dest.size, self.problem.n_nodes
        cost_a = None if cost is None else np.ascontiguousarray(cost, np.float64)
        mask = None if rt_allowed is None else np.ascontiguousarray(rt_allowed, np.uint8)
        lab = np.zeros((n, N)); pn = np.zeros((n, N), np.int32); pl = np.zeros((n, N), np.int32)
        ties = np.zeros(max(n, 1), np.int32)
        stats = np.zeros(6)
        self._ck(self.lib.dtl_backward_trees(self.h, tau, _ptr(cost_a), _ptr(mask), _ptr(dest), n, _ptr(lab), _ptr(pn), _ptr(pl),
                                             _ptr(ties), _ptr(stats)))
        st = dict(ms=stats[0], counted_edges=stats[1], relaxations=stats[2], pops=stats[3], rounds=stats[4],
                  replayed=stats[5])
        return lab, pn, pl, ties[:n], st

    # ---- ODME (Assignment::Demand_ODME, ODME.h:91-516) ----
    ODME_ROW = ("gap", "system_gap", "link_mae", "avg_tt", "ue_gap", "ue_gap_pct", "demand", "links_with_data")

    def odme_configure(self, at_pce, at_occ, sensors=()):
        a = np.ascontiguousarray(at_pce, np.float32); b = np.ascontiguousarray(at_occ, np.float32)
        tau = np.ascontiguousarray([s[0] for s in sensors] or [0], np.int32)
        link = np.ascontiguousarray([s[1] for s in sensors] or [0], np.int32)
        cnt = np.ascontiguousarray([s[2] for s in sensors] or [0], np.float32)
        ub = np.ascontiguousarray([s[3] for s in sensors] or [0], np.int32)
        self._ck(self.lib.dtl_odme_configure(self.h, _ptr(a), _ptr(b), len(sensors), _ptr(tau), _ptr(link), _ptr(cnt), _ptr(ub)))

    def odme_scatter(self, s: int) -> dict:
        row = np.zeros(8)
        self._ck(self.lib.dtl_odme_scatter(self.h, int(s), _ptr(row)))
        return dict(zip(self.ODME_ROW, row))

    def odme_update(self):
        self._ck(self.lib.dtl_odme_update(self.h))

    def odme_run(self, n: int):
        rows = np.zeros((n + 1, 8))
        done = C.c_int(0)
        self._ck(self.lib.dtl_odme_run(self.h, int(n), _ptr(rows), C.byref(done)))
        return rows, done.value

    def odme_state(self, tau: int = 0) -> dict:
        L = self.problem.n_links
        obs, ub, dev = np.zeros(L), np.zeros(L, np.int32), np.zeros(L)
        self._ck(self.lib.dtl_odme_get(self.h, tau, _ptr(obs), _ptr(ub), _ptr(dev)))
        return dict(obs_count=obs, upper_bound_flag=ub, est_count_dev=dev)

    # ---- sensitivity-analysis stage (main_api.cpp:771-1042) ----
    def sa_configure(self, rt_info, changes=(), dms=()):
        rt = np.ascontiguousarray(rt_info, np.int32)
        tau = np.ascontiguousarray([c[0] for c in changes] or [0], np.int32)
        link = np.ascontiguousarray([c[1] for c in changes] or [0], np.int32)
        lanes = np.ascontiguousarray([c[2] for c in changes] or [0.0], np.float64)
        self._ck(self.lib.dtl_sa_configure(self.h, _ptr(rt), len(changes), _ptr(tau), _ptr(link), _ptr(lanes)))
        if dms:  # "dms" rows of supply_side_scenario.csv: (period, link, information zone)
            d = np.ascontiguousarray(dms, np.int32).reshape(-1, 3)
            cols = [np.ascontiguousarray(d[:, i]) for i in range(3)]
            self.lib.dtl_sa_set_dms.restype = C.c_int
            self.lib.dtl_sa_set_dms.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            self._ck(self.lib.dtl_sa_set_dms(self.h, len(d), _ptr(cols[0]), _ptr(cols[1]), _ptr(cols[2])))

    def sa_route_modification(self) -> int:
        """g_column_pool_route_modification between the SA iterations and the SA column-pool iterations; returns the columns split"""
        n = C.c_int(0)
        self.lib.dtl_sa_route_modification.restype = C.c_int
        self.lib.dtl_sa_route_modification.argtypes = [C.c_void_p, C.c_void_p]
        self._ck(self.lib.dtl_sa_route_modification(self.h, C.byref(n)))
        return n.value

    def sa_begin(self) -> float:
        s = C.c_double(0)
        self._ck(self.lib.dtl_sa_begin(self.h, C.byref(s)))
        return s.value

    def sa_iteration(self, it: int) -> dict:
        row = np.zeros(2)
        self._ck(self.lib.dtl_sa_iteration(self.h, int(it), _ptr(row)))
        return dict(sysTT=row[0], avg_tt=row[1])

    def sa_column_update(self, n: int) -> dict:
        row = np.zeros(4)
        self._ck(self.lib.dtl_sa_column_update(self.h, int(n), _ptr(row)))
        return dict(avg_tt=row[0], total_gap=row[1], rel_gap_pct=row[2], demand=row[3])

    def sa_columns(self) -> dict:
        n = int(self.lib.dtl_num_columns(self.h))
        n_it = C.c_int(0)
        imp = np.zeros(max(n, 1), np.uint8)
        self._ck(self.lib.dtl_sa_get_columns(self.h, _ptr(imp), C.byref(n_it), None, None))
        ht = np.zeros((max(n_it.value, 1), max(n, 1))); hv = np.zeros((max(n_it.value, 1), max(n, 1)))
        self._ck(self.lib.dtl_sa_get_columns(self.h, None, None, _ptr(ht), _ptr(hv)))
        return dict(od_impact=imp[:n], time=ht[:n_it.value, :n], volume=hv[:n_it.value, :n])

    def vdf(self, tau: int, volume) -> dict:
        L = self.problem.n_links
        v = np.ascontiguousarray(volume, np.float64)
        names = ("tt", "doc", "voc", "P", "vt2")
        arrs = [np.zeros(L) for _ in names]
        self._ck(self.lib.dtl_vdf(self.h, tau, _ptr(v), *[_ptr(a) for a in arrs]))
        return dict(zip(names, arrs))

    def scatter(self, decay_k: int = -1):
        self._ck(self.lib.dtl_scatter(self.h, int(decay_k)))

    def tree(self, task: int):
        N = self.problem.n_nodes
        lab = np.zeros(N); pn = np.zeros(N, np.int32); pl = np.zeros(N, np.int32)
        self._ck(self.lib.dtl_get_tree(self.h, int(task), _ptr(lab), _ptr(pn), _ptr(pl)))
        return lab, pn, pl

    def timing(self, reset: bool = False) -> dict:
        ms = np.zeros(8); cnt = np.zeros(8, np.int64)
        self._ck(self.lib.dtl_get_timing(self.h, _ptr(ms), _ptr(cnt), int(reset)))
        return {"ms": dict(zip(self.TIMING_CLASSES, ms.tolist())), "launches": dict(zip(self.TIMING_CLASSES, cnt.tolist()))}

    def set_bounded_search(self, on: bool) -> bool:
        """Destination-bounded origin-bundle search inside the assignment loop (exact; on by default, False = the reference's complete trees); returns the old setting."""
        self.lib.dtl_set_bounded_search.restype = C.c_int
        self.lib.dtl_set_bounded_search.argtypes = [C.c_void_p, C.c_int]
        return bool(self.lib.dtl_set_bounded_search(self.h, int(bool(on))))

    def set_overlap(self, on: bool) -> bool:
        """Scatter of an iteration beside its shortest-path search (default) or everything on one stream; returns the old setting."""
        self.lib.dtl_set_overlap.restype = C.c_int
        self.lib.dtl_set_overlap.argtypes = [C.c_void_p, C.c_int]
        return bool(self.lib.dtl_set_overlap(self.h, int(bool(on))))

    def counters(self, reset: bool = False) -> dict:
        c = np.zeros(10)
        self._ck(self.lib.dtl_get_counters(self.h, _ptr(c), int(reset)))
        return dict(zip(self.COUNTERS, c.tolist()))

    K1_KERNELS = {0: "dtl::sssp_kernel", 1: "dtl::sssp_smq_kernel", 2: "dtl::sssp_bundle_kernel + dtl::bundle_finish_kernel",
                  3: "dtl::sssp_bundle_async_kernel + dtl::bundle_finish_kernel"}

    @property
    def sssp_kernel(self) -> str:
        """Name of the K1 kernel(s) the last shortest-path launch used."""
        self.lib.dtl_sssp_kernel_used.restype = C.c_int
        self.lib.dtl_sssp_kernel_used.argtypes = [C.c_void_p]
        return self.K1_KERNELS.get(int(self.lib.dtl_sssp_kernel_used(self.h)), "none")

    @property
    def device_bytes(self) -> int:
        return int(self.lib.dtl_device_bytes(self.h))
