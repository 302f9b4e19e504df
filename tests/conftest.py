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


@pytest.fixture(scope="session")
def problems(native):
    """GMNS fixtures read through the native reader (pinned against the reference in test_gmns_reader.py)."""
    import dlsim_mrm_b200 as D
    cache = {}

    def get(name, blocks=4):
        key = (name, blocks)
        if key not in cache:
            cache[key] = D.read_gmns(os.path.join(FIXTURES, name), blocks)
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
