import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def O():
    """the CPU oracle (test infrastructure)"""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session")
def ctx(pkg):
    """GPU context through the C ABI; fails loudly (no skip) when the library or the GPU is missing"""
    return pkg.default_context()


@pytest.fixture(scope="session")
def gold():
    out = {}
    for name in ("truth_mpmath", "truth_ld", "scipy_pairs", "philox_kat"):
        with open(os.path.join(GOLD, name + ".json")) as f:
            out[name] = json.load(f)
    return out


def relerr(a, b, floor=1e-3):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.abs(a - b) / np.maximum(floor, np.abs(b))
