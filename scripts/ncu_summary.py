"""Summarise an ncu capture for profiles/: key metrics of each kernel from `ncu -i X.ncu-rep --page raw --csv`
and, optionally, the hottest source lines from `--page source --csv`.

usage: python scripts/ncu_summary.py raw.csv [source.csv] > profiles/rNN_<what>.md
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of `{sys.argv[1]}`\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## {d.get('Kernel Name')}  (launch id {d.get('ID')})\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
        print()
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        per_fn, fn = {}, ""
        for r in rows:
            if r and r[0] == "Function Name":
                fn = r[1]
            elif r and r[0].strip().isdigit() and len(r) > 7 and r[4].isdigit():
                per_fn.setdefault(fn, []).append((int(r[4]), r[0], r[7], r[1]))
        for fn, lines in per_fn.items():
            tot = sum(x[0] for x in lines) or 1
            print(f"## hottest source lines of `{fn}` (warp stall samples, all)\n")
            print("| line | samples | share | instructions executed | source |\n|---|---|---|---|---|")
            for s_, ln, ie, src in sorted(lines, reverse=True)[:12]:
                print(f"| {ln} | {s_} | {100.0 * s_ / tot:.1f}% | {ie} | `{src.strip()[:100]}` |")
            print()


if __name__ == "__main__":
    main()
