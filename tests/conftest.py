import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def r_golden():
    return json.load(open(os.path.join(GOLDEN, "r_design_100x2.json")))


@pytest.fixture(scope="session")
def pymaxpro_golden():
    g = json.load(open(os.path.join(GOLDEN, "pymaxpro_vectors.json")))
    for c in g["cases"]:
        c["design"] = np.array([[float.fromhex(v) for v in row] for row in c["design_hex"]])
    return g


class _Switches:
    """Test switches of the library (maxpro_debug_set): force a kernel path for one test."""

    def __init__(self):
        self._touched = set()

    def set(self, name, value):
        from maxpro_b200 import _lib
        _lib.debug_set(name, value)
        self._touched.add(name)

    def clear(self, name):
        from maxpro_b200 import _lib
        _lib.debug_set(name, None)

    def reset(self):
        for name in self._touched:
            self.clear(name)
        self._touched.clear()


@pytest.fixture
def switches():
    s = _Switches()
    yield s
    s.reset()


def random_lhd(rng, n, d):
    """Random LHD of the reference's shape (src/lhd.rs:117-128), NumPy-generated."""
    x = np.empty((n, d))
    for k in range(d):
        x[:, k] = (rng.permutation(n) + rng.random(n)) / n
    return x


def lhd_bins_ok(x):
    """The reference's own validity test (src/lhd.rs:134-155): floor(x*n) hits every bin once."""
    n = x.shape[0]
    return all(sorted(np.floor(x[:, k] * n).astype(np.int64).tolist()) == list(range(n)) for k in range(x.shape[1]))
