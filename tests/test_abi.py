"""The C-ABI library loads and exports every symbol include/dtalite_b200.h declares; without a GPU every
compute entry point fails loudly (no CPU fallback).  CPU only."""
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dtalite_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(dtl_[a-z_0-9]+|network_assignment|dtalite_main)\s*\(", text)
    return sorted(set(names))


def test_header_symbols_are_exported(native):
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(native, n), f"{n} is declared in include/dtalite_b200.h but not exported"


def test_no_cpu_fallback(native, problems):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import dlsim_mrm_b200 as D
    with pytest.raises(D.EngineError, match="no CPU fallback"):
        D.Engine(problems("two_corridor"))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dlsim-mrm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle: the product path must not depend on it"


def test_every_included_header_is_a_build_dependency():
    """A header that only some .cu files include (e.g. csrc/*.cuh) must make the library stale when it changes: the build is
    in-tree and the .so travels to the GPU box as built here."""
    from dlsim_mrm_b200 import build
    deps = {os.path.realpath(d) for d in build.dependencies()}
    csrc = os.path.join(ROOT, "dlsim-mrm_b200", "csrc")
    inc = os.path.join(ROOT, "include")
    for src in build.dependencies():
        for name in re.findall(r'^\s*#\s*include\s+"([^"]+)"', open(src).read(), re.M):
            for base in (csrc, inc):
                path = os.path.realpath(os.path.join(base, name))
                if os.path.exists(path):
                    assert path in deps, f"{os.path.basename(src)} includes {name}, which the build does not watch"
