#!/usr/bin/env python
"""bench.py -- UE-assignment throughput of the B200 DTALite engine on the synthetic regional grid
(BASELINE.json configs[4]: 100k nodes / 300k links / 2000 zones / 4M trips).

A "step" is one column-generation iteration of the assignment loop (VDF -> volume scatter + MSA decay ->
all shortest-path trees -> tree back-trace / column update; reference main_api.cpp:589-728) over the whole
network.  `--gpus N` shards origin zones over N ranks (spatially contiguous blocks of zones), one process per GPU (torchrun);
the only data-path collective is the all-reduce of the fixed-point link-volume vector.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload grid100k] [--impl reference]

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own OpenMP CPU path
(oracle/_ref/dtalite_ref, the unmodified sources) on the same workload, all origin zones, for the same W + K iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ue_assignment_iterations_per_s"
UNIT = "iter/s"
BYTES_PER_EDGE = 20.0  # col_to 4 B + cost 8 B + label[to] 8 B per counted edge relaxation


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_dir(name: str) -> str:
    return os.path.join(tempfile.gettempdir(), f"dtl_bench_{name}_{os.getuid()}")


def ensure_workload(name: str) -> str:
    from dlsim_mrm_b200 import synth
    d = workload_dir(name)
    if not os.path.exists(os.path.join(d, ".complete")):
        tmp = d + f".tmp{os.getpid()}"
        shutil.rmtree(tmp, ignore_errors=True)
        synth.write_grid(tmp, **synth.WORKLOADS[name])
        open(os.path.join(tmp, ".complete"), "w").close()
        shutil.rmtree(d, ignore_errors=True)
        os.replace(tmp, d)
    return d


def k1_traffic(n_gpus: int, suffix: str = ""):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the same command
    (profiles/k1_traffic.json, written by scripts/ncu_traffic.py); None when no capture of this GPU count is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            t = json.load(f)
        return t.get(str(n_gpus) + suffix, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def other_kernels(tm, ctr, K, p, peak):
    """K2a / K3 / K4 of a step: algorithmic GB/s (12 B per back-traced path node; 12 B per column-link reference;
    160 B per (link, period)) over the event-timed phase, and the fraction of the measured HBM peak."""
    out = {}
    for name, phase, nbytes in (("K2a backtrace", "backtrace", 12.0 * ctr["path_nodes"]), ("K3 scatter", "scatter", 12.0 * ctr["link_refs"]),
                                ("K4 vdf", "vdf", 160.0 * float(p.n_links) * float(p.n_periods) * K)):
        ms = tm["ms"].get(phase, 0.0)
        if ms > 0:
            gbs = nbytes / (ms / 1e3) / 1e9
            out[name] = {"ms_per_step": ms / K, "achieved": gbs, "unit": "GB/s", "frac": gbs / peak}
    return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def reference_sample(workload: str, steps: int, warmup: int, n_origins: int | None = None):
    """Run the reference's own CPU path (oracle/_ref/dtalite_ref trace mode, unmodified sources) on the workload -- by default
    the FULL demand (every origin zone: nothing is extrapolated; the bound on the CPU time is the number of iterations the
    caller asks for), with `n_origins` the demand of the first S origin zones only.  Returns per-iteration seconds of the
    timed steps, the number of trees per iteration, the thread count and the sample description."""
    from oracle import binding as B
    if not os.path.exists(B.REF_BIN):
        raise FileNotFoundError(f"{B.REF_BIN} is missing (built by oracle/build_ref.sh where /root/reference exists)")
    src = ensure_workload(workload)
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    blocks = max(1, min(cores, 100))                     # number_of_memory_blocks = parallel width of the reference
    S = n_origins or full_trees(workload)
    with tempfile.TemporaryDirectory(prefix="dtl_ref_") as tmp:
        for f in ("node.csv", "link.csv", "settings.csv"):
            os.symlink(os.path.join(src, f), os.path.join(tmp, f))
        n_rows, trips = 0, 0.0
        with open(os.path.join(src, "demand.csv")) as fi, open(os.path.join(tmp, "demand.csv"), "w") as fo:
            fo.write(fi.readline())
            for line in fi:
                if int(line.split(",", 1)[0]) <= S:
                    fo.write(line)
                    n_rows += 1
                    trips += float(line.rsplit(",", 1)[1])
        t0 = time.time()
        _, tm = B.run_reference(tmp, "trace", warmup + steps, 0, blocks, "dump.bin", env={"DTL_TRACE_LIGHT": "1", "DTL_TRACE_NO_COLUMNS": "1"},
                                threads=blocks, timeout=3000)
        wall = time.time() - t0
    iters = tm["iter_s"][warmup:warmup + steps]
    trees = tm["n_sssp"] / tm["K"]
    if S >= full_trees(workload):
        sample = (f"{workload}: the whole workload ({n_rows} OD rows, {trees:.0f} shortest-path trees per iteration, nothing "
                  f"extrapolated), iterations {warmup}..{warmup + steps - 1} of a {warmup}+{steps}-iteration run of oracle/_ref/dtalite_ref "
                  f"(unmodified reference sources), OMP threads = memory blocks = {blocks}")
    else:
        sample = (f"{workload}: full network, demand of origin zones 1..{S} only ({n_rows} OD rows, {trees:.0f} shortest-path trees "
                  f"per iteration), {warmup}+{steps} iterations, OMP threads = memory blocks = {blocks}; iteration time scaled by "
                  f"(trees in the full workload / trees in the sample)")
    return dict(iter_s=iters, trees=trees, threads=int(tm["threads"]), blocks=blocks, sample=sample, wall_s=wall, io_s=tm["io_s"],
                counted_edges_total=tm["counted_edges"] * tm["K"], sssp_thread_s=tm["sssp_thread_s"], n_origins=S, od_rows=n_rows,
                trips=trips, nodes=tm["nodes"], links=tm["links"], zones=tm["zones"])


def full_trees(workload: str) -> int:
    from dlsim_mrm_b200 import synth
    return synth.WORKLOADS[workload]["n_zones"]  # 1 agent type x 1 period: one tree per zone


def run_reference_arm(args, rank):
    if rank != 0:
        return
    base = {"metric": METRIC, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    try:
        r = reference_sample(args.workload, args.steps, args.warmup)
    except Exception as ex:  # the oracle always exists in a complete checkout; say why if it does not
        print(json.dumps({**base, "unavailable": str(ex)[:300]}))
        return
    scale = full_trees(args.workload) / r["trees"]
    sec = statistics.mean(r["iter_s"]) * scale
    value = 1.0 / sec
    out = {**base, "value": value, "ms_per_step": sec * 1e3,
           # same workload as the engine arm (its `config` keys; nodes / links without the centroids and connectors both sides add)
           "config": {"workload": args.workload, "nodes": int(r["nodes"] - r["zones"]), "links": int(r["links"] - 2 * r["zones"]),
                      "zones": int(r["zones"]), "od_rows": int(r["od_rows"]), "trips": round(r["trips"], 1), "agent_types": 1, "periods": 1,
                      "first_timed_iteration": args.warmup, "sample": r["sample"]},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["threads"], "kind": "reference", "sample": r["sample"]},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "reference_io_s": r["io_s"], "reference_wall_s": r["wall_s"]}
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------------
def run_engine_arm(args, rank, world, local_rank):
    import torch
    import dlsim_mrm_b200 as D
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    from dlsim_mrm_b200 import sharding

    def max_over_ranks(x: float) -> float:
        return sharding.max_over_ranks(dist, x, device="cuda")

    if rank == 0:
        d = ensure_workload(args.workload)
    barrier()
    d = workload_dir(args.workload)
    if not os.path.exists(os.path.join(d, ".complete")):  # other node-local ranks wait for rank 0's files
        d = ensure_workload(args.workload)
    t0 = time.time()
    p = D.read_gmns(d, 8)
    read_s = time.time() - t0
    K, W = args.steps, args.warmup

    uid = None
    if world > 1:  # rank 0 makes the NCCL id once per process group; engines join (and share) that communicator
        uid = sharding.broadcast_bytes(dist, D.Engine.comm_unique_id() if rank == 0 else None, 128, device="cuda")

    def make_engine():
        # the dense-OD variant stores ~200 path links per column for 4 M OD cells: give the path-link arena what it needs
        arena = (40 << 30) // max(1, world) if args.workload.endswith("_dense") else 0
        e = D.Engine(p, device=local_rank, rank=rank, world_size=world, max_columns_per_od=K + W + 8,
                     delta_factor=args.delta, block_threads=args.block, ctas_per_sm=args.ctas, column_arena_bytes=arena)
        if world > 1:
            e.comm_init(uid)
        return e

    # ---- device-resident throughput: inputs already in HBM when the timed region starts ----
    e = make_engine()
    if not e.set_bounded_search(True):  # the engine's default (dtl_set_bounded_search): searches stop at the last destination with demand
        raise SystemExit("bench.py: the destination-bounded search is not the engine's default (DTL_COMPLETE_SEARCH set?)")
    e.iterations(0, W)
    e.timing(reset=True); e.counters(reset=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    t0 = time.perf_counter()
    rows = e.iterations(W, K)   # dtl_iterations: the K steps issued as a pipeline (back-trace of step k beside the search of step k+1)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    tm, ctr = e.timing(), e.counters()
    dev_s = max_over_ranks(tm["ms"]["iteration"] / 1e3)     # CUDA events on the engine's stream, max over ranks
    wall = max_over_ranks(wall)
    launches = int(sum(tm["launches"].values()))
    k1_kernel = e.sssp_kernel
    k1_kernels_per_launch = 2 if "bundle" in k1_kernel else 1  # the bundle variant is a search kernel + a label->tree kernel
    k1_launches = max(1, tm["launches"]["sssp"] // k1_kernels_per_launch)
    sssp_ms_launch = tm["ms"]["sssp"] / k1_launches       # event-timed duration of one K1 launch (all its kernels)
    edges_rank = rows[-1]["counted_edges"] / world  # the row holds the all-reduced count; ranks own equal shares of zones
    launches_per_iter = k1_launches / K
    peak, peak_src = measured_peak()
    sssp_ms_all = max_over_ranks(tm["ms"]["sssp"])
    dev_bytes = e.device_bytes
    # every rank must hold the same all-reduced integers, and they must be the single-GPU ones (committed table)
    import hashlib
    vol_sha = hashlib.sha256(e.volume_fixed(0).tobytes()).hexdigest()
    tag = float(int(vol_sha[:12], 16))
    same_on_all_ranks = max_over_ranks(tag) == tag and -max_over_ranks(-tag) == tag
    golden_sha = None
    try:
        with open(os.path.join(ROOT, "tests", "golden", "grid100k_volume_digests.json")) as f:
            golden_sha = json.load(f)["digests"].get(str(W + K)) if args.workload == "grid100k" else None
    except Exception:
        pass
    # the scatter runs beside K1 inside the timed steps; its duration ALONE (and that of the other kernels) comes from three more
    # iterations with every kernel on one stream
    e.set_overlap(False); e.timing(reset=True); e.counters(reset=True)
    for k in range(3):
        e.iteration(W + K + k)
    tm1, ctr1 = e.timing(), e.counters()
    k1_alone_ms = tm1["ms"]["sssp"] / max(1, tm1["launches"]["sssp"] // k1_kernels_per_launch)
    e.close()
    if not same_on_all_ranks:
        raise SystemExit(f"bench.py: rank {rank} holds link volumes {vol_sha[:16]}.. that differ from another rank's")
    if golden_sha is not None and golden_sha != vol_sha:
        raise SystemExit(f"bench.py: link volumes after {W + K} iterations ({vol_sha[:16]}..) differ from the single-GPU golden ({golden_sha[:16]}..)")

    # ---- end to end through the C ABI with host buffers: image upload + K iterations + result read-back ----
    barrier()
    t0 = time.perf_counter()
    e2 = make_engine()
    torch.cuda.synchronize()
    create_s = time.perf_counter() - t0
    e2.iterations(0, K)
    e2.finalize(K)
    st = e2.link_state(0, fields=("tt", "pce_volume"))   # the result of the assignment: link travel times and volumes
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = sum(a.nbytes for a in p.arrays.values())
    d2h = sum(a.nbytes for a in st.values()) + K * 16
    e2.close()
    barrier()

    # ---- the same steps with complete one-to-all trees (dtl_set_bounded_search(0): what the reference computes, routing.h:633-939).
    # Exact either way -- the link volumes must carry the same digest.  The complete run also gives the edge accounting of the
    # headline run: the Graph500 count (row[4]) is that of complete trees, a bounded search relaxes only a part of them, and the
    # kernels count their relaxations -- settled edges = complete count x (relaxations of the headline run / of the complete run).
    barrier()
    e3 = make_engine()
    e3.set_bounded_search(False)
    e3.iterations(0, W)
    e3.timing(reset=True); e3.counters(reset=True)
    barrier()
    e3.iterations(W, K)
    torch.cuda.synchronize()
    tm3, ctr3 = e3.timing(), e3.counters()
    complete_s = max_over_ranks(tm3["ms"]["iteration"] / 1e3)
    complete_sha = hashlib.sha256(e3.volume_fixed(0).tobytes()).hexdigest()
    complete_k1_ms = tm3["ms"]["sssp"] / max(1, tm3["launches"]["sssp"] // k1_kernels_per_launch)
    e3.close()
    if complete_sha != vol_sha:
        raise SystemExit(f"bench.py: complete and destination-bounded searches give different link volumes ({complete_sha[:16]}.. / {vol_sha[:16]}..)")
    settled_share = min(1.0, ctr["relaxations"] / max(1.0, ctr3["relaxations"]))
    pops_share = min(1.0, ctr["pops"] / max(1.0, ctr3["pops"]))  # the same ratio on node expansions (cross-check of the share above)
    settled_rank = edges_rank * settled_share
    achieved = BYTES_PER_EDGE * (settled_rank / launches_per_iter) / (sssp_ms_launch / 1e3) / 1e9
    gteps = rows[-1]["counted_edges"] * settled_share * K / (sssp_ms_all / 1e3) / 1e9
    complete_achieved = BYTES_PER_EDGE * (edges_rank / launches_per_iter) / (complete_k1_ms / 1e3) / 1e9

    out = None
    if rank == 0:
        value = K / dev_s
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_s / K * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "nodes": int(p.n_nodes - p.n_zones), "links": int(p.n_links - 2 * p.n_zones),
                       "zones": int(p.n_zones), "od_rows": int(p.n_od), "trips": float(p.total_demand_volume),
                       "agent_types": int(p.n_agent_types), "periods": int(p.n_periods),
                       "parallelism": f"origin zones sharded over {world} GPU(s) in spatially contiguous blocks", "first_timed_iteration": W,
                       "l2": "inputs larger than L2 (trees 2.4 GB + column pool per iteration vs 126 MB L2)"},
            "clocks": clocks,
            "e2e": {"value": K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": d2h // K,
                    "what": "dtl_create from host arrays + K iterations + finalize + link-state read-back, wall clock"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": k1_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": k1_traffic(world), "peak_source": peak_src,
                         "algorithmic_bytes_per_edge": BYTES_PER_EDGE, "counted_edges_per_launch": settled_rank / launches_per_iter,
                         "edge_count": "edges a destination-bounded launch settles = Graph500 count of the complete trees "
                                       f"({edges_rank / launches_per_iter:.0f}) x {settled_share:.4f}, the kernel-counted relaxations of the timed "
                                       f"steps over those of the same steps with complete trees (node expansions: x {pops_share:.4f})",
                         "kernel_ms_per_launch": sssp_ms_launch,
                         # the same kernel without the scatter beside it (the three one-stream iterations behind the timed steps)
                         "kernel_ms_per_launch_alone": k1_alone_ms,
                         "frac_alone": BYTES_PER_EDGE * (settled_rank / launches_per_iter) / (k1_alone_ms / 1e3) / 1e9 / peak if k1_alone_ms else None},
            "sssp_gteps": gteps,
            # the other kernels of a step against the same HBM peak (algorithmic bytes of SURVEY.md 8d; this rank's share)
            "other_kernels": other_kernels(tm1, ctr1, 3, p, peak),
            "other_kernels_note": "three extra iterations with every kernel on one stream (inside the timed steps the scatter and its "
                                  "all-reduce run on a second stream beside K1: phase_ms_per_step['scatter'] is that overlapped duration)",
            "one_stream_ms_per_step": tm1["ms"]["iteration"] / 3,
            "complete_trees": {"value": K / complete_s, "unit": UNIT, "ms_per_step": complete_s / K * 1e3, "k1_ms_per_launch": complete_k1_ms,
                               "roofline_achieved": complete_achieved, "roofline_frac": complete_achieved / peak,
                               "counted_edges_per_launch": edges_rank / launches_per_iter, "traffic": k1_traffic(world, "_complete"),
                               "same_link_volume_digest": True,
                               "what": "the same K steps with dtl_set_bounded_search(0): every search builds the reference's complete "
                                       "one-to-all tree (routing.h:633-939).  The headline `value` runs the engine's default, where a search "
                                       "stops when nothing queued can still improve a destination with demand -- same columns, volumes, gaps "
                                       "(the digest is checked in this run)"},
            "wall_ms_per_step": wall / K * 1e3,
            "phase_ms_per_step": {k_: v / K for k_, v in tm["ms"].items()},
            "work_per_step": {k_: v / K for k_, v in ctr.items() if k_ != "counted_edges"},
            "counted_edges_per_step": rows[-1]["counted_edges"] * settled_share,
            "complete_tree_edges_per_step": rows[-1]["counted_edges"],
            "relative_gap_last": rows[-1]["gap"],
            "volume_digest": {"sha256": vol_sha, "iterations": W + K, "same_on_all_ranks": bool(same_on_all_ranks),
                              "matches_single_gpu_golden": None if golden_sha is None else golden_sha == vol_sha},
            "device_gib": dev_bytes / 2**30, "read_gmns_s": read_s, "e2e_create_s": create_s, "e2e_total_s": e2e_s,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = reference_sample(args.workload, args.cpu_steps, 1)
                scale = full_trees(args.workload) / r["trees"]
                sec = statistics.mean(r["iter_s"]) * scale
                out["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": r["threads"], "kind": "reference",
                                       "sample": r["sample"],
                                       "sssp_gteps_per_core": r["counted_edges_total"] / max(1e-9, r["sssp_thread_s"]) / 1e9}
            except Exception as ex:
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {ex}"[:300]}
        print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="grid100k")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--delta", type=float, default=0.0)
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.warmup < 3 and args.impl == "engine":
        log("bench.py: note: fewer than 3 warm-up steps")
    if args.impl == "reference":
        run_reference_arm(args, rank)
    else:
        run_engine_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
