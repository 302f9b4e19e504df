"""Array image of a GMNS network + demand, the data contract between the host side and the engine.

This is the SoA re-expression of the reference's AoS data model (``CLink`` / ``CPeriod_VDF`` /
``CNode`` / ``COZone`` / the 4-D column pool header, reference ``src/cpp/DTA.h:518-714,1095-1452``,
``VDF.h:369-466``), laid out exactly as SURVEY.md section 9 "Graph image" describes:

* nodes: physical nodes in ``node.csv`` order, then one centroid per zone in ascending ``zone_id``;
* links: ``link.csv`` order, then two connectors per activity node (``main_api.cpp:534-570``);
* per-period arrays are ``[T, L]``, per-period-per-agent-type arrays ``[T, AT, L]`` (C order);
* demand is the list of OD cells with ``od_volume > 0`` sorted by ``(o, d, at, tau)``.

The class only holds and validates arrays; it computes nothing.
"""
from __future__ import annotations

from dataclasses import dataclass, field, fields
from typing import Dict

import numpy as np

# (name, dtype, shape code) -- shape codes: N nodes, L links, Z zones, A agent types,
# TL = [T, L], TAL = [T, AT, L], OD = [n_od]
_ARRAYS = [
    ("link_from", np.int32, "L"), ("link_to", np.int32, "L"), ("link_tag", np.int32, "L"),
    ("node_zone", np.int32, "N"), ("zone_centroid", np.int32, "Z"),
    ("fftt", np.float64, "TL"), ("cap", np.float64, "TL"), ("alpha", np.float64, "TL"), ("beta", np.float64, "TL"),
    ("vf", np.float64, "TL"), ("vc", np.float64, "TL"), ("nlanes", np.float64, "TL"), ("cap_lane", np.float64, "TL"),
    ("qdf", np.float64, "TL"), ("preload", np.float64, "TL"), ("penalty", np.float64, "TL"),
    ("Q_alpha", np.float64, "TL"), ("Q_beta", np.float64, "TL"), ("Q_cd", np.float64, "TL"), ("Q_n", np.float64, "TL"),
    ("Q_cp", np.float64, "TL"), ("Q_s", np.float64, "TL"), ("t2", np.float64, "TL"), ("L_h", np.float64, "TL"),
    ("start_h", np.float64, "TL"), ("end_h", np.float64, "TL"), ("plf", np.float64, "TL"),
    ("cycle_length", np.float32, "TL"), ("red_time", np.float32, "TL"), ("sat_flow", np.float32, "TL"),
    ("vdf_type", np.int32, "TL"),
    ("pce", np.float64, "TAL"), ("occ", np.float64, "TAL"), ("toll", np.float64, "TAL"), ("lr_price", np.float64, "TAL"),
    ("allowed", np.uint8, "TAL"),
    ("vot", np.float32, "A"),
    ("od_o", np.int32, "OD"), ("od_d", np.int32, "OD"), ("od_at", np.int32, "OD"), ("od_tau", np.int32, "OD"),
    ("od_vol", np.float64, "OD"),
    ("origin_demand", np.float32, "Z"),
]
ARRAY_SPECS = {n: (dt, sh) for n, dt, sh in _ARRAYS}


@dataclass
class Problem:
    n_nodes: int
    n_links: int
    n_zones: int
    n_agent_types: int
    n_periods: int
    memory_blocks: int = 4          # number_of_memory_blocks (reference default, flash_dta.cpp:212-218)
    total_demand_volume: float = 0.0  # float32 accumulator of the reference (DTA.h, Assignment)
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)
    # optional identifiers kept for writers / debugging (not used by the engine)
    node_id: np.ndarray | None = None
    zone_id: np.ndarray | None = None
    # sensitivity-analysis inputs (not part of dtl_problem): real_time_info per agent type, "sa" lane changes (tau, link, lanes)
    rt_info: np.ndarray | None = None
    sa_changes: list | None = None
    dms_links: list | None = None   # "dms" rows of supply_side_scenario.csv: (period, link, information zone seq no or -1)
    # ODME inputs: agent-type PCE / occupancy as floats, measurement.csv link rows (tau, link, count, upper_bound_flag)
    at_pce: np.ndarray | None = None
    at_occ: np.ndarray | None = None
    sensors: list | None = None

    def __getattr__(self, name):
        arrs = self.__dict__.get("arrays", {})
        if name in arrs:
            return arrs[name]
        raise AttributeError(name)

    @property
    def n_od(self) -> int:
        return int(self.arrays["od_o"].shape[0])

    def shape_of(self, code: str):
        N, L, Z, A, T = self.n_nodes, self.n_links, self.n_zones, self.n_agent_types, self.n_periods
        return {"N": (N,), "L": (L,), "Z": (Z,), "A": (A,), "TL": (T, L), "TAL": (T, A, L),
                "OD": (self.arrays["od_o"].shape[0],) if "od_o" in self.arrays else (0,)}[code]

    def finalize(self) -> "Problem":
        """Coerce dtypes/contiguity and check shapes and the OD sort order."""
        for name, (dt, code) in ARRAY_SPECS.items():
            if name not in self.arrays:
                raise ValueError(f"Problem is missing array {name!r}")
            a = np.ascontiguousarray(self.arrays[name], dtype=dt)
            want = self.shape_of(code)
            if a.shape != want:
                a = a.reshape(want)
            self.arrays[name] = a
        o, d, at, tau = (self.arrays[k].astype(np.int64) for k in ("od_o", "od_d", "od_at", "od_tau"))
        if o.size:
            key = ((o * self.n_zones + d) * self.n_agent_types + at) * self.n_periods + tau
            if np.any(np.diff(key) <= 0):
                raise ValueError("OD cells must be unique and sorted by (o, d, at, tau)")
            if not np.all(self.arrays["od_vol"] > 0):
                raise ValueError("OD cells must have od_volume > 0")
        return self

    def with_defaults(self) -> "Problem":
        """Fill every array that is absent with the reference's constructor defaults
        (CPeriod_VDF(), VDF.h:57-74; link reader defaults, input.h:3494-3520)."""
        T, A, L = self.n_periods, self.n_agent_types, self.n_links
        dflt = dict(alpha=0.15, beta=4.0, preload=0.0, penalty=0.0, Q_alpha=0.272876961, Q_beta=4.0,
                    Q_cd=0.954946463, Q_n=1.141574427, Q_cp=0.400089684, Q_s=4.0, plf=1.0,
                    cycle_length=-1.0, red_time=0.0, sat_flow=1800.0, vdf_type=1,
                    pce=1.0, occ=1.0, toll=0.0, lr_price=0.0, allowed=1)
        for name, (dt, code) in ARRAY_SPECS.items():
            if name not in self.arrays and name in dflt:
                self.arrays[name] = np.full(self.shape_of(code), dflt[name], dtype=dt)
        return self


def subset_origins(p: Problem, zones) -> Problem:
    """Copy of ``p`` whose demand keeps only the given origin zones (used for bounded samples)."""
    zones = np.asarray(sorted(set(int(z) for z in zones)), dtype=np.int32)
    keep = np.isin(p.arrays["od_o"], zones)
    arrs = dict(p.arrays)
    for k in ("od_o", "od_d", "od_at", "od_tau", "od_vol"):
        arrs[k] = p.arrays[k][keep].copy()
    od = np.zeros(p.n_zones, dtype=np.float32)
    mask = np.zeros(p.n_zones, dtype=bool)
    mask[zones] = True
    od[mask] = p.arrays["origin_demand"][mask]
    arrs["origin_demand"] = od
    q = Problem(p.n_nodes, p.n_links, p.n_zones, p.n_agent_types, p.n_periods, p.memory_blocks,
                float(np.float32(arrs["od_vol"].astype(np.float32).sum())) if keep.any() else 0.0, arrs,
                p.node_id, p.zone_id)
    return q.finalize()
