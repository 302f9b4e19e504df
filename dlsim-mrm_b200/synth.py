"""Seeded synthetic regional grid in GMNS files (BASELINE.json config 5).

``write_grid(dir, rows, cols, n_links, n_zones, ...)`` writes node.csv, link.csv, demand.csv and settings.csv
that both the reference executable and this engine read.  Defaults give the named size:
316 x 317 = 100 172 nodes, 300 000 directed links, 2 000 zones, 4 000 000 trips kept on each origin's 200
heaviest destinations (400 000 OD rows).  Link lengths are jittered per directed link so that no two paths
tie exactly in FP64 (exact ties are an artefact of uniform grids).
"""
from __future__ import annotations

import os

import numpy as np

SEED = 20240601

SETTINGS = """[assignment],,,column_generation_iterations,column_updating_iterations,sensitivity_analysis_iterations,simulation_iterations,number_of_memory_blocks
,,,{K},0,0,0,{blocks}
,,,,,,,
[agent_type],agent_type,name,,vot,flow_type,pce,
,p,passenger,,10,0,1,
,,,,,,,
[link_type],link_type,link_type_name,,agent_type_blocklist,type_code,traffic_flow_code,
,1,freeway,,,f,0,
,2,arterial,,,a,0,
,3,local,,,a,0,
,,,,,,,
[demand_period],demand_period_id,demand_period,,time_period,,,
,1,AM,,0700_0800,,,
,,,,,,,
[demand_file_list],file_sequence_no,file_name,,format_type,demand_period,agent_type,
,1,demand.csv,,column,AM,p,
,,,,,,,
[output_file_configuration],,path_output,major_path_volume_threshold,trajectory_sampling_rate,trajectory_diversion_only,td_link_performance_sampling_interval_in_min,
,,1,0.1,1,0,1,
"""


def _spanning_tree_mask(n_nodes, eu, ev, rng):
    """Random spanning tree (Kruskal on shuffled undirected edges); returns a mask over the undirected edges."""
    parent = np.arange(n_nodes)

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    order = rng.permutation(eu.size)
    keep = np.zeros(eu.size, bool)
    left = n_nodes - 1
    for i in order:
        a, b = find(int(eu[i])), find(int(ev[i]))
        if a != b:
            parent[a] = b
            keep[i] = True
            left -= 1
            if left == 0:
                break
    return keep


def make_grid(rows=316, cols=317, n_links=300_000, n_zones=2000, total_trips=4_000_000.0, dest_per_origin=200, seed=SEED):
    """Arrays of the synthetic network (no files)."""
    rng = np.random.default_rng(seed)
    N = rows * cols
    r, c = np.divmod(np.arange(N), cols)
    # undirected lattice edges
    h = np.flatnonzero(c + 1 < cols)
    v = np.flatnonzero(r + 1 < rows)
    eu = np.concatenate([h, v])
    ev = np.concatenate([h + 1, v + cols])
    tree = _spanning_tree_mask(N, eu, ev, rng)          # both directions of the tree keep the graph strongly connected
    directed_u = np.concatenate([eu, ev])
    directed_v = np.concatenate([ev, eu])
    must = np.concatenate([tree, tree])
    n_total = directed_u.size
    n_links = int(min(max(n_links, int(must.sum())), n_total))
    extra = rng.permutation(np.flatnonzero(~must))[: n_links - int(must.sum())]
    keep = must.copy()
    keep[extra] = True
    idx = np.flatnonzero(keep)
    idx = idx[rng.permutation(idx.size)]                # link.csv order is arbitrary, like real GMNS exports
    lu, lv = directed_u[idx], directed_v[idx]
    L = lu.size
    # hierarchy: a link is on row/col line `k` if both ends share that row (horizontal) or column (vertical)
    ru, cu = np.divmod(lu, cols)
    rv, cv = np.divmod(lv, cols)
    line = np.where(ru == rv, ru, cu)
    freeway = line % 32 == 0
    arterial = (line % 8 == 0) & ~freeway
    local_fast = rng.random(L) < 0.5
    free_speed = np.where(freeway, 65.0, np.where(arterial, 45.0, np.where(local_fast, 35.0, 25.0)))
    lanes = np.where(freeway, 3, np.where(arterial, 2, 1))
    capacity = np.where(freeway, 1900.0, np.where(arterial, 1400.0, 900.0))
    link_type = np.where(freeway, 1, np.where(arterial, 2, 3))
    length = np.round(rng.uniform(200.0, 1200.0, L), 4)  # metres, jittered per directed link
    x = -112.0 + c * 0.0070 + rng.uniform(-0.001, 0.001, N)
    y = 33.0 + r * 0.0060 + rng.uniform(-0.001, 0.001, N)
    # zones: one activity node per cell of a jittered sub-lattice
    side = int(np.ceil(np.sqrt(n_zones)))
    zr = ((np.arange(side) + 0.5) * rows / side).astype(int)
    zc = ((np.arange(side) + 0.5) * cols / side).astype(int)
    cells = [(a, b) for a in zr for b in zc]
    sel = rng.permutation(len(cells))[:n_zones]
    zone_nodes = []
    used = set()
    for i in sorted(sel):
        a = int(np.clip(cells[i][0] + rng.integers(-2, 3), 0, rows - 1))
        b = int(np.clip(cells[i][1] + rng.integers(-2, 3), 0, cols - 1))
        n = a * cols + b
        while n in used:
            n = (n + 1) % N
        used.add(n)
        zone_nodes.append(n)
    zone_nodes = np.array(zone_nodes)
    Z = zone_nodes.size
    node_zone = np.zeros(N, np.int64)
    node_zone[zone_nodes] = np.arange(1, Z + 1)
    # gravity demand, each origin's heaviest destinations only
    P = rng.uniform(0.5, 1.5, Z)
    A = rng.uniform(0.5, 1.5, Z)
    zx, zy = c[zone_nodes] * 0.7, r[zone_nodes] * 0.7   # km on a ~700 m lattice
    dkm = np.hypot(zx[:, None] - zx[None, :], zy[:, None] - zy[None, :])
    Tm = P[:, None] * A[None, :] * np.exp(-0.08 * dkm)
    np.fill_diagonal(Tm, 0.0)
    kdest = int(min(dest_per_origin, Z - 1))
    part = np.argpartition(-Tm, kdest - 1, axis=1)[:, :kdest]
    mask = np.zeros_like(Tm, bool)
    np.put_along_axis(mask, part, True, axis=1)
    Tm = np.where(mask, Tm, 0.0)
    Tm *= total_trips / Tm.sum()
    o, d = np.nonzero(Tm)
    vol = np.round(Tm[o, d], 4)
    ok = vol > 0
    return dict(N=N, x=x, y=y, node_zone=node_zone, link_from=lu, link_to=lv, length=length, free_speed=free_speed, lanes=lanes,
                capacity=capacity, link_type=link_type, od_o=o[ok] + 1, od_d=d[ok] + 1, od_vol=vol[ok], n_zones=Z)


def write_grid(directory, K=20, blocks=8, **kw):
    """Write the GMNS files of a synthetic grid into ``directory``; returns the array dict of ``make_grid``."""
    g = make_grid(**kw)
    os.makedirs(directory, exist_ok=True)
    N = g["N"]
    zid = g["node_zone"]
    with open(os.path.join(directory, "node.csv"), "w") as f:
        f.write("node_id,x_coord,y_coord,zone_id\n")
        ztxt = np.where(zid > 0, zid.astype(str), "")
        f.write("\n".join(f"{i + 1},{g['x'][i]:.6f},{g['y'][i]:.6f},{ztxt[i]}" for i in range(N)))
        f.write("\n")
    with open(os.path.join(directory, "link.csv"), "w") as f:
        f.write("link_id,from_node_id,to_node_id,link_type,length,lanes,free_speed,capacity\n")
        lf, lt = g["link_from"] + 1, g["link_to"] + 1
        f.write("\n".join(f"{i + 1},{lf[i]},{lt[i]},{g['link_type'][i]},{g['length'][i]:.4f},{g['lanes'][i]},"
                          f"{g['free_speed'][i]:.0f},{g['capacity'][i]:.0f}" for i in range(lf.size)))
        f.write("\n")
    with open(os.path.join(directory, "demand.csv"), "w") as f:
        f.write("o_zone_id,d_zone_id,volume\n")
        f.write("\n".join(f"{a},{b},{v:.4f}" for a, b, v in zip(g["od_o"], g["od_d"], g["od_vol"])))
        f.write("\n")
    with open(os.path.join(directory, "settings.csv"), "w") as f:
        f.write(SETTINGS.format(K=K, blocks=blocks))
    return g


# named workloads
WORKLOADS = {
    # BASELINE.json configs[4]: 100k nodes / 300k links / 2000 zones / 4M trips
    "grid100k": dict(rows=316, cols=317, n_links=300_000, n_zones=2000, total_trips=4_000_000.0, dest_per_origin=200),
    # the dense-OD variant of SURVEY.md 8d: every zone pair carries demand (3 998 000 OD rows); GPU only -- the reference's
    # column pool would need ~40 GB of host memory
    "grid100k_dense": dict(rows=316, cols=317, n_links=300_000, n_zones=2000, total_trips=4_000_000.0, dest_per_origin=1999),
    # same generator, small enough for a CPU run in seconds (tests)
    "grid2k": dict(rows=45, cols=46, n_links=6_200, n_zones=60, total_trips=60_000.0, dest_per_origin=25),
    "grid10k": dict(rows=100, cols=101, n_links=30_000, n_zones=200, total_trips=400_000.0, dest_per_origin=60),
}

if __name__ == "__main__":
    import sys
    name, out = sys.argv[1], sys.argv[2]
    g = write_grid(out, **WORKLOADS[name])
    print(name, "nodes", g["N"], "links", g["link_from"].size, "zones", g["n_zones"], "od rows", g["od_o"].size)
