"""Build recipe of the native library (hand-written sm_100a CUDA + the C-ABI host layer).

``python -m dlsim_mrm_b200.build`` compiles ``csrc/*.cu`` / ``csrc/*.cpp`` into
``lib/libdtalite_b200.so`` in-tree (the built file is git-ignored but travels to the GPU box).
nvcc cross-compiles for sm_100a without a GPU.  ``-fmad=false`` keeps every FP64 multiply and add
individually rounded, like the reference's baseline x86-64 build, so parity can be bit-level.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdtalite_b200.so")
EXE_PATH = os.path.join(LIB_DIR, "DTALite_b200")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-Wall,-Wno-unused-function", "-I", INCLUDE]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the native library cannot be built")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name: str, sssp_flags, verbose: bool = False) -> str:
    """Experiment helper: a copy of the library with extra nvcc flags on kernels_*.cu -> lib/variants/<name>.so
    (selected at run time with DTL_LIB=<path>)."""
    vdir = os.path.join(LIB_DIR, "variants")
    odir = os.path.join(vdir, "obj_" + name)
    os.makedirs(odir, exist_ok=True)
    objs = []
    for s in [x for x in sources() if os.path.basename(x) != "main.cpp"]:
        o = os.path.join(odir, os.path.basename(s) + ".o")
        extra = list(sssp_flags) if os.path.basename(s).startswith("kernels_") else []
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose and extra else []) + \
            ["-x", "cu" if s.endswith(".cu") else "c++", "-c", s, "-o", o]
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if out.returncode != 0:
            raise RuntimeError(out.stdout.decode())
        if verbose and extra:
            sys.stderr.write("\n".join(l for l in out.stdout.decode().splitlines() if "registers" in l or "sssp_kernel" in l) + "\n")
        objs.append(o)
    lib = os.path.join(vdir, name + ".so")
    subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs + ["-ldl"])
    return lib


def dependencies():
    """Every file the library is built from: sources plus the headers they include (.h and .cuh of csrc/, include/*.h)."""
    return (sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh"))
            + glob.glob(os.path.join(INCLUDE, "*.h")))


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile the shared library (and the DTALite_b200 executable) if sources changed."""
    os.makedirs(LIB_DIR, exist_ok=True)
    deps = dependencies()
    lib_src = [s for s in sources() if os.path.basename(s) != "main.cpp"]
    if force or _stale(LIB_PATH, deps):
        objs = []
        procs = []
        objdir = os.path.join(LIB_DIR, "obj")
        os.makedirs(objdir, exist_ok=True)
        for s in lib_src:
            o = os.path.join(objdir, os.path.basename(s) + ".o")
            objs.append(o)
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu" if s.endswith(".cu") else "c++",
                                                                                      "-c", s, "-o", o]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        for s, p in procs:
            out = p.communicate()[0].decode()
            if p.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}:\n{out}")
            if verbose or "warning" in out:
                sys.stderr.write(out)
        subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs + ["-ldl"])
    main_src = os.path.join(CSRC, "main.cpp")
    if os.path.exists(main_src) and (force or _stale(EXE_PATH, deps + [LIB_PATH])):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", INCLUDE, main_src, "-o", EXE_PATH,
                               "-L", LIB_DIR, "-ldtalite_b200", "-Wl,-rpath,$ORIGIN"])
    return LIB_PATH


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
