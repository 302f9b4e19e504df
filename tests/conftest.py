import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FIXTURES = os.path.join(ROOT, "tests", "fixtures")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def native():
    """The built native library (host-side symbols work without a GPU)."""
    from dlsim_mrm_b200 import build, engine
    build.build_native()
    return engine.load_library()


def stage_fixture(name, dst):
    """tests/fixtures/<base> copied into dst, then the files of tests/fixtures/overlays/<name> for names like "<base>+odme" """
    import shutil
    base = name.split("+")[0]
    for f in os.listdir(os.path.join(FIXTURES, base)):
        shutil.copy(os.path.join(FIXTURES, base, f), dst)
    if "+" in name:
        ov = os.path.join(FIXTURES, "overlays", name)
        for f in os.listdir(ov):
            shutil.copy(os.path.join(ov, f), dst)


@pytest.fixture(scope="session")
def problems(native, tmp_path_factory):
    """GMNS fixtures read through the native reader (pinned against the reference in test_gmns_reader.py)."""
    import dlsim_mrm_b200 as D
    cache = {}

    def get(name, blocks=4):
        key = (name, blocks)
        if key not in cache:
            d = os.path.join(FIXTURES, name)
            if "+" in name:
                d = str(tmp_path_factory.mktemp("fixture"))
                stage_fixture(name, d)
            cache[key] = D.read_gmns(d, blocks)
        return cache[key]
    return get


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLDEN, f"{name}_ref.npz"))
        return cache[name]
    return get


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import binding
    binding.build_oracle()
    return binding


def rel_err(a, b, floor=1e-9):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
