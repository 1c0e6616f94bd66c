"""pytest configuration: the `gpu` marker, repo-root imports, shared helpers.

`-m "not gpu"` : oracle vs reference / golden vectors, host graph builder, C-ABI surface (no device work).
`-m gpu`       : parity tests proper -- the CUDA path through the C ABI against the oracle and the goldens.
"""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def upper_from_full(cab):
    Q = cab.shape[0]
    return [cab[a, b] for a in range(Q) for b in range(a, Q)]


def rel_err(a, b, floor=1e-300):
    """max elementwise |a-b| / max(|b|, floor)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


@pytest.fixture(scope="session")
def built():
    """The product library and the checkers, built in-tree (no-op when already built)."""
    import __graft_entry__ as ge
    ge.build()
    return True
