"""Write profiles/k1_traffic.json from an `ncu --set full` raw CSV of the bench command's SSSP launch:
dram__bytes_read.sum + dram__bytes_write.sum of that launch (what bench.py reports as roofline.traffic).
usage: python scripts/ncu_traffic.py <raw.csv> <n_gpus> <note>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
out_path = os.path.join(ROOT, "profiles", "k1_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
tot, names, grids, dur = 0.0, [], [], 0.0
for r in rows[2:]:  # one K1 launch: the single-origin kernels are one kernel, the origin-bundle variant is search + label->tree pass
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "")
    if "sssp" not in name and "bundle_finish" not in name:
        continue
    if names and not ("bundle" in names[0] and "bundle_finish" in name and len(names) == 1):
        break
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(d[k]) * UNIT[units[hdr.index(k)]]
    names.append(name); grids.append(d.get("launch__grid_size")); dur += float(d["gpu__time_duration.sum"])
    if "bundle" not in names[0]:
        break
if names:
    out[sys.argv[2]] = {"dram_bytes_per_launch": tot, "kernel": " + ".join(names), "grid": " + ".join(str(g) for g in grids),
                        "duration_ms_under_ncu": dur, "note": sys.argv[3] if len(sys.argv) > 3 else ""}
json.dump(out, open(out_path, "w"), indent=1)
print(out)
