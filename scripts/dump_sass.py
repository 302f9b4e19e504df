"""SASS listing of every kernel family of the library (the instantiation the engine launches by default) -> profiles/sass/<tag>_<kernel>.sass
plus an index with the instruction mix that matters for this path (global / shared atomics, reductions, barriers, 16-byte loads).
usage: python scripts/dump_sass.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dlsim-mrm_b200", "lib", "libdtalite_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out_dir = os.path.join(ROOT, "profiles", "sass")
os.makedirs(out_dir, exist_ok=True)
# template kernels: the instantiation that runs by default (engine.cu: make_bundle_launch / choose_sssp_variant)
PREFER = {
    # (the destination-bounded instantiation is the default inside the loop, the complete one runs the count pass and dtl_set_bounded_search(0))
    "sssp_bundle_async_kernel": ("<768, 16, 4, false, 1, true>", "<768, 16, 4, false, 1, false>"), "sssp_bundle_kernel": "<1024, 16, 2>", "bundle_finish_kernel": "<16, 64, false>",
    "bundle_pred_kernel": "<16>", "bundle_pred_helper_kernel": "<768, 16>", "sssp_kernel": "<128, 4>", "sssp_smq_kernel": ("<512, 2, 2, true>", "<512, 2, 2, false>"), "scatter_kernel": None, "vdf_kernel": None,
}
txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, check=True).stdout.decode()
blocks = re.split(r"\n\s*Function : ", txt)[1:]
demangle = lambda m: subprocess.run(["c++filt", m], stdout=subprocess.PIPE).stdout.decode().strip()
index = []
for b in blocks:
    mangled = b.split("\n", 1)[0].strip()
    name = demangle(mangled)
    short = re.sub(r"^void ", "", name)
    short = re.sub(r"\(.*$", "", short).replace("dtl::", "")
    family = short.split("<")[0]
    targs = short[len(family):]
    want = PREFER.get(family)
    if want is not None and targs.replace("(bool)", "").replace("(int)", "") not in ((want,) if isinstance(want, str) else want):
        continue
    lines = [l for l in b.split("\n")[1:] if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    ops = collections.Counter()
    for l in lines:
        m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops[m.group(1)] += 1
    fname = f"{tag}_{family}{re.sub(r'[^0-9A-Za-z]+', '_', targs).strip('_')}.sass"
    with open(os.path.join(out_dir, fname), "w") as f:
        f.write(f"// {name}\n// cuobjdump -sass of dlsim-mrm_b200/lib/libdtalite_b200.so (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false)\n")
        f.write("\n".join(lines) + "\n")
    grp = lambda pat: sum(v for k, v in ops.items() if re.match(pat, k))
    index.append((short, fname, len(lines), grp(r"ATOMG"), grp(r"REDG|RED\b|RED\."), grp(r"ATOMS"), grp(r"BAR|WARPSYNC"), grp(r"LDG.*128"),
                  grp(r"LDG"), grp(r"STG"), grp(r"SHFL"), grp(r"DADD|DMUL|DFMA|DSETP|MUFU"), grp(r"UCGABAR|UTMA|UTCMMA|TCGEN|UBLKCP")))
with open(os.path.join(out_dir, f"{tag}_INDEX.md"), "w") as f:
    f.write(f"# SASS listings ({tag}) -- one file per kernel family, the instantiation the engine launches by default\n\n")
    f.write("Columns: static instruction counts.  `cluster/TMA/tcgen05` = UCGABAR / UTMA* / UBLKCP / UTCMMA: none of the kernels uses thread-block\n"
            "clusters, the tensor memory accelerator or the tensor cores -- the path is scattered 8- and 16-byte gather/atomic work (DESIGN.md 3).\n\n")
    f.write("| kernel | file | instr | ATOMG | REDG | ATOMS | BAR/WARPSYNC | LDG.128 | LDG | STG | SHFL | FP64 | cluster/TMA/tcgen05 |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for row in sorted(index):
        f.write("| " + " | ".join(str(x) for x in row) + " |\n")
print(len(index), "kernels")
