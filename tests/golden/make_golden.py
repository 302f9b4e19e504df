"""Regenerates the golden fixtures under tests/golden/ from the reference itself.  Run in the build
container (needs /root/reference and oracle/_ref/dtalite_ref, built by oracle/build_ref.sh):

    python tests/golden/make_golden.py

Outputs (committed):
  *_final_summary.txt   the line ranges of the reference's own committed final_summary.csv that pin this
                        path: two-corridor lines 35-58, 4_101_corridor lines 150-177
  *_ref.npz             state dumped from the reference engine (oracle/_ref/dtalite_ref trace mode) on
                        tests/fixtures/<name>: per-iteration rows, link states, the column pool, and
                        shortest-path trees (full arrays for small cases, SHA-256 digests otherwise)
"""
import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as B  # noqa: E402

REF = os.environ.get("DTL_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def copy_lines(src, first, last, dst):
    with open(src) as f:
        lines = f.readlines()
    with open(dst, "w") as f:
        f.write(f"# lines {first}-{last} of {os.path.relpath(src, REF)} (reference repository), verbatim\n")
        f.writelines(lines[first - 1:last])


def run_case(name, K, CU, tree_iters, full_tree_zones, multi_period=False):
    src = os.path.join(ROOT, "tests", "fixtures", name)
    with tempfile.TemporaryDirectory() as tmp:
        for f in os.listdir(src):
            shutil.copy(os.path.join(src, f), tmp)
        env = {"DTL_TRACE_TREES": ",".join(str(k) for k in tree_iters)}
        B.run_reference(tmp, "trace", K, CU, 4, "dump.bin", env=env, threads=4, multi_period=multi_period)
        d = B.load_dump(os.path.join(tmp, "dump.bin"))
    N, L, Z, AT, T = (int(x) for x in d["meta"])
    out = {"meta": d["meta"], "K": np.int32(K), "CU": np.int32(CU)}
    out["iter_rows"] = np.stack([d[f"iter_row:k={k}"] for k in range(K)])  # k, sysTT, least, gap, avg_tt
    for key in ("tt", "pce_volume", "person_volume", "vdf_link_volume", "doc", "voc", "P", "vt2"):
        out[f"final_{key}"] = d[f"{key}:final:tau=0"]
        if key in ("tt", "pce_volume"):
            for k in (2, K - 1):
                out[f"k{k}_{key}_after_vdf"] = d[f"{key}:after_vdf:k={k}:tau=0"]
            for n in range(CU):
                out[f"cu{n}_{key}"] = d[f"{key}:cu={n}:tau=0"]
    for tau in range(1, T):  # further demand periods: same keys with a ":tau=t" suffix
        for key in ("tt", "pce_volume", "person_volume", "vdf_link_volume", "doc", "voc", "P", "vt2"):
            out[f"final_{key}:tau={tau}"] = d[f"{key}:final:tau={tau}"]
            if key in ("tt", "pce_volume"):
                for k in (2, K - 1):
                    out[f"k{k}_{key}_after_vdf:tau={tau}"] = d[f"{key}:after_vdf:k={k}:tau={tau}"]
                for n in range(CU):
                    out[f"cu{n}_{key}:tau={tau}"] = d[f"{key}:cu={n}:tau={tau}"]
    out["final_sysTT"] = d["final_sysTT"]
    for k in tree_iters:
        for at in range(AT):
            out[f"gen_cost_k{k}_at{at}"] = d[f"gen_cost:k={k}:at={at}:tau=0"]
            for tau in range(1, T):
                out[f"gen_cost_k{k}_at{at}_tau{tau}"] = d[f"gen_cost:k={k}:at={at}:tau={tau}"]
    for key in ("col_o", "col_d", "col_at", "col_tau", "col_node_sum", "col_seq", "col_nlinks", "col_nnodes", "col_volume",
                "col_cost", "col_time"):
        out[key] = d[key + ":final"]
    links = d["col_links:final"]
    if links.size <= 200000:
        out["col_links"] = links
    out["col_links_sha256"] = np.array(sha(links))
    # per-column digest so a mismatch can be located without shipping every link id
    off = np.concatenate([[0], np.cumsum(d["col_nlinks:final"])]).astype(np.int64)
    w = (np.arange(links.size, dtype=np.int64) - np.repeat(off[:-1], d["col_nlinks:final"]) + 1)
    out["col_link_checksum"] = np.add.reduceat((links.astype(np.int64) + 1) * w, off[:-1]) if links.size else np.zeros(0, np.int64)
    tree_keys, tree_sha = [], []
    for key in sorted(k for k in d if k.startswith("label:")):
        tag = key[len("label:"):]
        z = int(tag.split("z=")[1])
        tree_keys.append(tag)
        tree_sha.append(sha(d["label:" + tag]) + sha(d["pred_link:" + tag]) + sha(d["pred_node:" + tag]))
        if full_tree_zones is None or z in full_tree_zones:
            out["tree_label:" + tag] = d["label:" + tag]
            out["tree_pred_link:" + tag] = d["pred_link:" + tag]
            out["tree_pred_node:" + tag] = d["pred_node:" + tag]
    out["tree_keys"] = np.array(tree_keys)
    out["tree_sha256"] = np.array(tree_sha)
    # network image digests pin the native GMNS reader without the reference binary
    p = B.problem_from_dump(d)
    names = sorted(p.arrays)
    out["problem_array_names"] = np.array(names)
    out["problem_array_sha256"] = np.array([sha(p.arrays[n]) for n in names])
    out["total_demand_volume"] = np.float32(p.total_demand_volume)
    np.savez_compressed(os.path.join(OUT, f"{name}_ref.npz"), **out)
    print(name, "->", os.path.getsize(os.path.join(OUT, f"{name}_ref.npz")) // 1024, "KiB")


def run_backward(cases):
    """backward trees (optimal_backward_label_correcting_from_destination, routing.h:1009) of the reference on the final travel
    times of a K-iteration run -> backward_ref.npz: the cost array, SHA-256 of every dumped tree, full arrays of a few"""
    out = {}
    for name, K, stride, full in cases:
        src = os.path.join(ROOT, "tests", "fixtures", name)
        with tempfile.TemporaryDirectory() as tmp:
            for f in os.listdir(src):
                shutil.copy(os.path.join(src, f), tmp)
            B.run_reference(tmp, "trace", K, 0, 4, "dump.bin", env={"DTL_TRACE_BACKWARD": str(stride), "DTL_TRACE_NO_COLUMNS": "1"}, threads=4)
            d = B.load_dump(os.path.join(tmp, "dump.bin"))
        out[f"{name}:K"] = np.int32(K)
        out[f"{name}:cost"] = d["tt:final:tau=0"]  # avg_travel_time after the final VDF pass; RT_waiting_time is 0 without the simulator
        keys, shas = [], []
        for key in sorted(k for k in d if k.startswith("bw_label:")):
            tag = key[len("bw_label:"):]
            keys.append(tag)
            shas.append(sha(d["bw_label:" + tag]) + sha(d["bw_pred_link:" + tag]) + sha(d["bw_pred_node:" + tag]))
            if int(tag.split("z=")[1]) in full and ":at=0:" in tag:
                for part in ("label", "pred_link", "pred_node"):
                    out[f"{name}:{part}:{tag}"] = d[f"bw_{part}:{tag}"]
        out[f"{name}:keys"] = np.array(keys)
        out[f"{name}:sha256"] = np.array(shas)
    np.savez_compressed(os.path.join(OUT, "backward_ref.npz"), **out)
    print("backward trees ->", os.path.getsize(os.path.join(OUT, "backward_ref.npz")) // 1024, "KiB")


def stage_fixture(name, dst):
    """copies tests/fixtures/<base> into dst, then the files of tests/fixtures/overlays/<name> for names like "<base>+odme" """
    base = name.split("+")[0]
    for f in os.listdir(os.path.join(ROOT, "tests", "fixtures", base)):
        shutil.copy(os.path.join(ROOT, "tests", "fixtures", base, f), dst)
    if "+" in name:
        ov = os.path.join(ROOT, "tests", "fixtures", "overlays", name)
        for f in os.listdir(ov):
            shutil.copy(os.path.join(ov, f), dst)


def run_odme(cases):
    """OD demand estimation (Assignment::Demand_ODME, ODME.h:91-516) followed by the SA stage with 0 iterations: the state the
    UNMODIFIED network_assignment(1, K, CU, ODME, 0, 0, 4) leaves (harness mode golden_dump, ONE OpenMP thread: the reference's
    ODME scatter sums link volumes inside a parallel loop, so only a single-thread run is reproducible to the bit)"""
    out = {}
    for name, K, CU, n_odme in cases:
        key = f"{name}:ODME={n_odme}"
        with tempfile.TemporaryDirectory() as tmp:
            stage_fixture(name, tmp)
            B.run_reference(tmp, "golden_dump", K, CU, 4, "dump.bin", threads=1, extra_args=["0", str(n_odme)])
            d = B.load_dump(os.path.join(tmp, "dump.bin"))
            small = os.path.getsize(os.path.join(tmp, "route_assignment.csv")) < (1 << 20)
            for f in ("link_performance", "route_assignment"):
                if small:
                    shutil.copy(os.path.join(tmp, f + ".csv"), os.path.join(OUT, "outputs", f"{name}_odme{n_odme}_{f}.csv"))
            with open(os.path.join(tmp, "final_summary.csv")) as f, open(os.path.join(OUT, "outputs", f"{name}_odme{n_odme}_final_summary_rows.txt"), "w") as o:
                for line in f:
                    if line.startswith("ODME"):
                        o.write(line)
        out[f"{key}:K_CU_ODME"] = np.array([K, CU, n_odme], np.int32)
        for k2 in ("obs_count", "upper_bound_flag", "est_count_dev"):
            out[f"{key}:{k2}"] = d[f"{k2}:tau=0"]
        out[f"{key}:agent_pce"] = d["agent_pce"]; out[f"{key}:agent_occ"] = d["agent_occ"]
        for k2 in ("tt", "pce_volume", "person_volume"):
            out[f"{key}:final_{k2}"] = d[f"{k2}:final:tau=0"]
        for k2 in ("col_node_sum", "col_seq", "col_volume", "col_cost", "col_time"):
            out[f"{key}:{k2}"] = d[k2 + ":final"]
        out[f"{key}:col_links_sha256"] = np.array(sha(d["col_links:final"]))
    np.savez_compressed(os.path.join(OUT, "odme_ref.npz"), **out)
    print("ODME ->", os.path.getsize(os.path.join(OUT, "odme_ref.npz")) // 1024, "KiB")


def run_sa(cases):
    """sensitivity-analysis stage (main_api.cpp:771-1042): the state the UNMODIFIED network_assignment(1, K, CU, 0, S, 0, 4)
    leaves in the reference's globals (harness mode golden_dump) -> sa_ref.npz; its output files -> outputs/<name>_sa<S>_*.csv"""
    out = {}
    os.makedirs(os.path.join(OUT, "outputs"), exist_ok=True)
    for name, K, CU, S in cases:
        key = f"{name}:S={S}"
        with tempfile.TemporaryDirectory() as tmp:
            stage_fixture(name, tmp)
            # dms rows: g_column_pool_route_modification accumulates path travel times on columns that several origins share inside an
            # OpenMP loop over origins (assignment.h:1081,1219) -- only a single-thread run is reproducible
            B.run_reference(tmp, "golden_dump", K, CU, 4, "dump.bin", threads=1 if "+dms" in name else 4, extra_args=[str(S)])
            d = B.load_dump(os.path.join(tmp, "dump.bin"))
            for f in ("link_performance", "route_assignment"):
                if name.startswith("corridor_101") and S > 0:
                    continue  # 30 MB of text per run: the state of this case is pinned through sa_ref.npz only
                dst = os.path.join(OUT, "outputs", f"{name}_sa{S}_{f}.csv")
                shutil.copy(os.path.join(tmp, f + ".csv"), dst)
                if os.path.getsize(dst) > (1 << 20):  # large files are committed gzip-compressed
                    import gzip
                    with open(dst, "rb") as fi, gzip.GzipFile(dst + ".gz", "wb", 9, mtime=0) as fo:
                        shutil.copyfileobj(fi, fo)
                    os.remove(dst)
            with open(os.path.join(tmp, "final_summary.csv")) as f, open(os.path.join(OUT, "outputs", f"{name}_sa{S}_final_summary_rows.txt"), "w") as o:
                seen = False
                for line in f:
                    seen = seen or line.startswith("Iteration, CPU Running Time")
                    if seen and (line[:1].isdigit() and line.count(",") in (2, 4) or line.startswith("column updating: iteration")
                                 or line.startswith("Sensitivity analysis stage")):
                        o.write(line)
        out[f"{key}:K_CU_S"] = np.array([K, CU, S], np.int32)
        out[f"{key}:rt_info"] = d["agent_rt_info"]
        for k2 in ("design_flag", "sa_lanes", "sa_nlanes_after", "sa_cap_after", "sa_fftt_after"):
            out[f"{key}:{k2}"] = d[f"{k2}:tau=0"]
        for k2 in ("tt", "pce_volume", "person_volume", "doc", "P"):
            out[f"{key}:final_{k2}"] = d[f"{k2}:final:tau=0"]
        for k2 in ("col_o", "col_d", "col_at", "col_node_sum", "col_seq", "col_nlinks", "col_volume", "col_cost", "col_time",
                   "col_impacted_path_flag", "col_sa_flag", "col_od_impact_flag", "col_at_od_impacted"):
            out[f"{key}:{k2}"] = d[k2 + ":final"]
        links = d["col_links:final"]
        if links.size <= 100000:
            out[f"{key}:col_links"] = links
        out[f"{key}:col_links_sha256"] = np.array(sha(links))
    np.savez_compressed(os.path.join(OUT, "sa_ref.npz"), **out)
    print("sensitivity analysis ->", os.path.getsize(os.path.join(OUT, "sa_ref.npz")) // 1024, "KiB")


def reference_output_files(name, K, CU, multi_period=False):
    """link_performance.csv / route_assignment.csv as the reference's own network_assignment() writes them
    (harness mode "golden": the unmodified entry point, main_api.cpp:499) -> tests/golden/outputs/."""
    src = os.path.join(ROOT, "tests", "fixtures", name)
    os.makedirs(os.path.join(OUT, "outputs"), exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for f in os.listdir(src):
            shutil.copy(os.path.join(src, f), tmp)
        B.run_reference(tmp, "golden", K, CU, 4, None, threads=4, multi_period=multi_period)
        for f in ("link_performance", "route_assignment"):
            shutil.copy(os.path.join(tmp, f + ".csv"), os.path.join(OUT, "outputs", f"{name}_{f}.csv"))
        # final_summary.csv: the rows this path produces (iteration rows main_api.cpp:724-727, column-pool rows assignment.h:909-911)
        with open(os.path.join(tmp, "final_summary.csv")) as f, open(os.path.join(OUT, "outputs", f"{name}_final_summary_rows.txt"), "w") as o:
            seen = False
            for line in f:
                seen = seen or line.startswith("Iteration, CPU Running Time")
                if seen and (line[:1].isdigit() and line.count(",") == 4 or line.startswith("column updating: iteration")):
                    o.write(line)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "odme":
        run_odme((("sa_corridor+odme", 20, 5, 12), ("corridor_101+odme", 20, 5, 20)))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sa":
        run_sa((("sa_corridor", 20, 5, 3), ("sa_corridor", 20, 5, 0), ("corridor_101", 20, 5, 0), ("corridor_101", 20, 5, 2),
                ("corridor_101+dms", 20, 5, 2), ("corridor_101+dms", 20, 5, 0)))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "backward":
        run_backward((("two_corridor", 20, 1, (0, 1)), ("edge_cases", 10, 1, (0, 3)), ("corridor_101", 20, 3, (0, 45)),
                      ("toll_signal", 10, 2, (2,))))
        sys.exit(0)
    for name_, K_, CU_, mp_ in (("two_corridor", 20, 0, False), ("corridor_3", 20, 20, False), ("edge_cases", 10, 4, False),
                                ("corridor_3_2p", 12, 6, True), ("toll_signal", 10, 4, False), ("asu_qvdf", 20, 3, False)):
        reference_output_files(name_, K_, CU_, mp_)
    copy_lines(os.path.join(REF, "src/test_case_01_two_corridor/final_summary.csv"), 35, 58,
               os.path.join(OUT, "two_corridor_final_summary.txt"))
    copy_lines(os.path.join(REF, "dataset/4_101_corridor/final_summary.csv"), 150, 177,
               os.path.join(OUT, "corridor_101_final_summary.txt"))
    run_case("two_corridor", 20, 0, range(20), None)
    run_case("corridor_101", 20, 5, (0, 2, 19), {7})
    run_case("corridor_3", 20, 20, range(20), None)       # authored fixture, BASELINE.json configs[1]
    run_case("asu_qvdf", 20, 3, (0, 2, 19), None)         # authored fixture, queue VDF, BASELINE.json configs[2]
    # two demand periods + allowed_uses filter; needs the MAX_TIMEPERIODS = 4 build of the reference (oracle/build_ref.sh)
    run_case("corridor_3_2p", 12, 6, range(12), None, multi_period=True)
    run_case("edge_cases", 10, 4, range(10), None)        # authored corner cases: unreachable zones, closed links, skipped demand rows
    run_case("toll_signal", 10, 4, range(10), None)       # authored: tolls, penalties, values of time, signalised links
