"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle.so (the CPU restatement) and a
loader for the tagged binary dumps written by oracle/_ref/dtalite_ref (the reference engine itself).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (dlsim-mrm_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "dtalite_ref")
REF_BIN_TP4 = os.path.join(HERE, "_ref", "dtalite_ref_tp4")  # same sources with MAX_TIMEPERIODS = 4 (multi-period fixtures)

_DT = {0: np.int32, 1: np.float64, 2: np.float32, 3: np.int64}


def build_oracle(force: bool = False) -> str:
    """gcc-compile the C restatement (seconds)."""
    src = os.path.join(HERE, "dtalite_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "dtalite_oracle.h"))):
        subprocess.check_call(["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w",
                               "-o", LIB_PATH, src, "-lm"])
    return LIB_PATH


def load_dump(path: str) -> dict:
    """Read a dtalite_ref trace file: {name: ndarray}."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    n = len(data)
    while pos < n:
        (ln,) = struct.unpack_from("<I", data, pos); pos += 4
        name = data[pos:pos + ln].decode(); pos += ln
        dt = data[pos]; pos += 1
        (cnt,) = struct.unpack_from("<Q", data, pos); pos += 8
        dtype = np.dtype(_DT[dt])
        out[name] = np.frombuffer(data, dtype=dtype, count=cnt, offset=pos).copy()
        pos += cnt * dtype.itemsize
    return out


def problem_from_dump(d: dict, memory_blocks: int = 4):
    """Build the engine's Problem image from the reference's own in-memory network (dump_network)."""
    import importlib
    Problem = importlib.import_module("dlsim_mrm_b200.problem").Problem
    N, L, Z, AT, T = (int(x) for x in d["meta"])
    arrs = {}
    arrs["link_from"] = d["link_from"]; arrs["link_to"] = d["link_to"]; arrs["link_tag"] = d["link_tag"]
    zid2seq = {int(z): i for i, z in enumerate(d["zone_id"])}
    nz = np.full(N, -1, dtype=np.int32)
    for i, zid in enumerate(d["node_zone_id"]):
        if zid >= 1:
            nz[i] = zid2seq[int(zid)]
    arrs["node_zone"] = nz
    arrs["zone_centroid"] = d["zone_node"]
    per_tau = dict(fftt="vdf_fftt", cap="vdf_cap", alpha="vdf_alpha", beta="vdf_beta", vf="vdf_vf", vc="vdf_vc",
                   nlanes="vdf_nlanes", cap_lane="vdf_cap_lane", qdf="vdf_qdf", preload="vdf_preload",
                   penalty="vdf_penalty", Q_alpha="vdf_Q_alpha", Q_beta="vdf_Q_beta", Q_cd="vdf_Q_cd", Q_n="vdf_Q_n",
                   Q_cp="vdf_Q_cp", Q_s="vdf_Q_s", t2="vdf_t2", L_h="vdf_L", start_h="vdf_start_h",
                   end_h="vdf_end_h", plf="vdf_plf", cycle_length="vdf_cycle", red_time="vdf_red",
                   sat_flow="vdf_sat", vdf_type="vdf_type")
    for k, src in per_tau.items():
        arrs[k] = np.stack([d[f"{src}:tau={t}"] for t in range(T)])
    per_at = dict(pce="vdf_pce", occ="vdf_occ", toll="vdf_toll", lr_price="vdf_lr_price", allowed="link_allowed")
    for k, src in per_at.items():
        arrs[k] = np.stack([np.stack([d[f"{src}:tau={t}:at={a}"] for a in range(AT)]) for t in range(T)])
    arrs["vot"] = d["agent_vot"].astype(np.float32)
    for k in ("od_o", "od_d", "od_at", "od_tau"):
        arrs[k] = d[k]
    arrs["od_vol"] = d["od_volume"]
    arrs["origin_demand"] = d["origin_demand"].astype(np.float32)
    p = Problem(N, L, Z, AT, T, memory_blocks, float(np.float32(d["total_demand_volume"][0])), arrs,
                node_id=d["node_id"], zone_id=d["zone_id"])
    return p.finalize()


# --------------------------------------------------------------------------------------------
class _OrcProblem(C.Structure):
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


class _VdfOut(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("tt", "doc", "voc", "P", "vt2", "avg_queue_speed", "avg_speed_bpr", "Q_mu",
                                          "Q_gamma", "t0", "t3", "severe_congestion_P", "lane_based_D",
                                          "avg_waiting_time")]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Stateful wrapper: one instance == one reference run on one Problem."""

    def __init__(self, problem):
        build_oracle()
        self.lib = C.CDLL(LIB_PATH)
        L = self.lib
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_void_p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_num_tasks.argtypes = [C.c_void_p]
        L.orc_get_tasks.argtypes = [C.c_void_p] * 4
        L.orc_set_tree_capture.argtypes = [C.c_void_p] * 4
        L.orc_iteration.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_column_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_finalize.restype = C.c_double
        L.orc_finalize.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_link_state.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 8
        L.orc_get_model_speed.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_get_gen_cost.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_num_columns.restype = C.c_longlong
        L.orc_num_columns.argtypes = [C.c_void_p]
        L.orc_num_column_links.restype = C.c_longlong
        L.orc_num_column_links.argtypes = [C.c_void_p]
        L.orc_get_columns.argtypes = [C.c_void_p] * 10
        L.orc_forward_star.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.orc_sssp.restype = C.c_longlong
        L.orc_sssp.argtypes = [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 4
        L.orc_vdf_link.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_sa_configure.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_sa_lane_change.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_sa_begin.restype = C.c_double
        L.orc_sa_begin.argtypes = [C.c_void_p]
        L.orc_sa_iteration.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_sa_column_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_odme_configure.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_odme_sensor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int]
        L.orc_odme_scatter.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_odme_update.argtypes = [C.c_void_p]
        L.orc_backward_sssp.restype = C.c_longlong
        L.orc_backward_sssp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.problem = problem
        self._keep = dict(problem.arrays)
        cp = _OrcProblem()
        for n in ("n_nodes", "n_links", "n_zones", "n_agent_types", "n_periods", "memory_blocks"):
            setattr(cp, n, int(getattr(problem, n)))
        for name, _ in _OrcProblem._fields_:
            if name in self._keep:
                setattr(cp, name, _ptr(self._keep[name]))
        cp.n_od = problem.n_od
        cp.total_demand_volume = float(problem.total_demand_volume)
        self._cp = cp
        self.h = L.orc_create(C.byref(cp))
        self.n_tasks = L.orc_num_tasks(self.h)
        at = np.zeros(self.n_tasks, np.int32); tau = np.zeros_like(at); zone = np.zeros_like(at)
        L.orc_get_tasks(self.h, _ptr(at), _ptr(tau), _ptr(zone))
        self.task_at, self.task_tau, self.task_zone = at, tau, zone
        self._cap = None

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def capture_trees(self, on: bool = True):
        N = self.problem.n_nodes
        if on:
            self._cap = (np.zeros((self.n_tasks, N), np.float64), np.zeros((self.n_tasks, N), np.int32),
                         np.zeros((self.n_tasks, N), np.int32))
            self.lib.orc_set_tree_capture(self.h, *[_ptr(a) for a in self._cap])
        else:
            self._cap = None
            self.lib.orc_set_tree_capture(self.h, None, None, None)
        return self._cap

    def iteration(self, k: int):
        row = np.zeros(5)
        self.lib.orc_iteration(self.h, k, _ptr(row))
        return dict(sysTT=row[0], least=row[1], gap=row[2], avg_tt=row[3], counted_edges=int(row[4]))

    def column_update(self, n: int):
        row = np.zeros(4)
        self.lib.orc_column_update(self.h, n, _ptr(row))
        return dict(avg_tt=row[0], total_gap=row[1], rel_gap_pct=row[2], demand=row[3])

    def finalize(self, K: int) -> float:
        return float(self.lib.orc_finalize(self.h, K))

    def link_state(self, tau: int = 0):
        L = self.problem.n_links
        names = ("tt", "pce_volume", "person_volume", "link_volume", "doc", "voc", "P", "vt2")
        arrs = [np.zeros(L) for _ in names]
        self.lib.orc_get_link_state(self.h, tau, *[_ptr(a) for a in arrs])
        return dict(zip(names, arrs))

    def gen_cost(self, at: int, tau: int = 0):
        c = np.zeros(self.problem.n_links)
        self.lib.orc_get_gen_cost(self.h, at, tau, _ptr(c))
        return c

    def columns(self):
        n = self.lib.orc_num_columns(self.h)
        nl = self.lib.orc_num_column_links(self.h)
        i32 = lambda m: np.zeros(max(int(m), 1), np.int32)
        f64 = lambda m: np.zeros(max(int(m), 1), np.float64)
        od, ns, sq, nlk, nnd, vol, cost, tm, links = i32(n), i32(n), i32(n), i32(n), i32(n), f64(n), f64(n), f64(n), i32(nl)
        self.lib.orc_get_columns(self.h, *[_ptr(a) for a in (od, ns, sq, nlk, nnd, vol, cost, tm, links)])
        return dict(od_index=od[:n], node_sum=ns[:n], seq_no=sq[:n], n_links=nlk[:n], n_nodes=nnd[:n], volume=vol[:n],
                    cost=cost[:n], time=tm[:n], links=links[:nl])

    # ---- stand-alone kernels ----
    def forward_star(self, at: int, tau: int = 0):
        N, L = self.problem.n_nodes, self.problem.n_links
        rp = np.zeros(N + 1, np.int32); cl = np.zeros(max(L, 1), np.int32); ct = np.zeros(max(L, 1), np.int32)
        ne = C.c_int(0)
        self.lib.orc_forward_star(C.byref(self._cp), at, tau, _ptr(rp), _ptr(cl), _ptr(ct), C.byref(ne))
        return rp, cl[:ne.value].copy(), ct[:ne.value].copy()

    def sssp(self, fs, cost, origin_zone: int):
        rp, cl, ct = fs
        N = self.problem.n_nodes
        cost = np.ascontiguousarray(cost, np.float64)
        label = np.zeros(N); pn = np.zeros(N, np.int32); pl = np.zeros(N, np.int32)
        relax = C.c_longlong(0)
        cl_ = cl if cl.size else np.zeros(1, np.int32)
        ct_ = ct if ct.size else np.zeros(1, np.int32)
        pops = self.lib.orc_sssp(C.byref(self._cp), _ptr(rp), _ptr(cl_), _ptr(ct_), _ptr(cost), int(origin_zone),
                                 _ptr(label), _ptr(pn), _ptr(pl), C.byref(relax))
        return label, pn, pl, int(pops), int(relax.value)

    # ---- sensitivity-analysis stage (main_api.cpp:771-1042) ----
    def sa_configure(self, rt_info, changes=(), dms=()):
        rt = np.ascontiguousarray(rt_info, np.int32)
        self.lib.orc_sa_configure(self.h, _ptr(rt))
        for tau, link, lanes in changes:
            self.lib.orc_sa_lane_change(self.h, int(tau), int(link), float(lanes))
        self.lib.orc_sa_dms_link.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        for tau, link, zone in dms or ():  # "dms" rows: the reader only keeps links whose to-node carries a zone id
            self.lib.orc_sa_dms_link(self.h, int(tau), int(link), 1, int(zone))

    def sa_route_modification(self) -> int:
        """g_column_pool_route_modification (assignment.h:1077-1383), between the SA iterations and the SA column-pool iterations;
        returns the number of columns that were split (-1: the reference would have stopped)"""
        self.lib.orc_sa_route_modification.restype = C.c_int
        self.lib.orc_sa_route_modification.argtypes = [C.c_void_p]
        return int(self.lib.orc_sa_route_modification(self.h))

    def sa_begin(self) -> float:
        return float(self.lib.orc_sa_begin(self.h))

    def sa_iteration(self, it: int):
        row = np.zeros(3)
        self.lib.orc_sa_iteration(self.h, int(it), _ptr(row))
        return dict(sysTT=row[0], avg_tt=row[1], least=row[2])

    def sa_column_update(self, n: int):
        row = np.zeros(4)
        self.lib.orc_sa_column_update(self.h, int(n), _ptr(row))
        return dict(avg_tt=row[0], total_gap=row[1], rel_gap_pct=row[2], demand=row[3])

    # ---- ODME (ODME.h:91-516) ----
    ODME_ROW = ("gap", "system_gap", "link_mae", "avg_tt", "ue_gap", "ue_gap_pct", "demand", "links_with_data")

    def odme_configure(self, at_pce, at_occ, sensors=()):
        a = np.ascontiguousarray(at_pce, np.float32); b = np.ascontiguousarray(at_occ, np.float32)
        self.lib.orc_odme_configure(self.h, _ptr(a), _ptr(b))
        for tau, link, count, ub in sensors:
            self.lib.orc_odme_sensor(self.h, int(tau), int(link), float(count), int(ub))

    def odme_scatter(self, s: int):
        row = np.zeros(8)
        self.lib.orc_odme_scatter(self.h, int(s), _ptr(row))
        return dict(zip(self.ODME_ROW, row))

    def odme_update(self):
        self.lib.orc_odme_update(self.h)

    def backward_sssp(self, cost, dest_node: int, rt_allowed=None):
        N = self.problem.n_nodes
        cost = np.ascontiguousarray(cost, np.float64)
        mask = None if rt_allowed is None else np.ascontiguousarray(rt_allowed, np.uint8)
        label = np.zeros(N); pn = np.zeros(N, np.int32); pl = np.zeros(N, np.int32)
        pops = self.lib.orc_backward_sssp(C.byref(self._cp), _ptr(cost), None if mask is None else _ptr(mask), int(dest_node), _ptr(label), _ptr(pn), _ptr(pl))
        return label, pn, pl, int(pops)

    def vdf_link(self, tau: int, link: int, volume: float, want_speed: bool = False):
        out = _VdfOut()
        sp = np.zeros(300, np.float32) if want_speed else None
        self.lib.orc_vdf_link(C.byref(self._cp), tau, link, float(volume), C.byref(out), _ptr(sp) if want_speed else None)
        res = {n: getattr(out, n) for n, _ in _VdfOut._fields_}
        if want_speed:
            res["model_speed"] = sp
        return res


def run_reference(workdir: str, mode: str, K: int, CU: int, B: int, out: str | None = None, env: dict | None = None,
                  timeout: float = 3600, threads: int | None = None, multi_period: bool = False, extra_args=None):
    """Run oracle/_ref/dtalite_ref in ``workdir`` (stdin closed: the reference blocks on getchar()
    when it stops on an error, utils.cpp:41-46).  Returns (stdout_text, timing_dict_or_None)."""
    import json
    ref_bin = REF_BIN_TP4 if multi_period else REF_BIN
    if not os.path.exists(ref_bin):
        raise FileNotFoundError(ref_bin)
    e = dict(os.environ)
    if threads:
        e["OMP_NUM_THREADS"] = str(threads)
    if env:
        e.update(env)
    cmd = [ref_bin, mode, str(K), str(CU), str(B)] + ([out] if out else []) + list(extra_args or [])
    r = subprocess.run(cmd, cwd=workdir, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       env=e, timeout=timeout)
    text = r.stdout.decode(errors="replace")
    if "Program stops" in text:
        raise RuntimeError("reference stopped on an input error:\n" + text[-2000:])
    if r.returncode != 0:
        raise RuntimeError(f"reference exited {r.returncode}:\n" + text[-2000:])
    timing = None
    for line in text.splitlines():
        if line.startswith("REFTIMING "):
            timing = json.loads(line[len("REFTIMING "):])
    return text, timing
