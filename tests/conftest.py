import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import isr_oracle as O
    O.lib()
    return O


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def assert_matrix_parity(G, Gref, Eref=None, rel=1e-9, abs_floor=2e-12, what="matrix", row_rel=0.0):
    """|G - Gref| <= rel*|Gref| + abs_floor + Eref (+ row_rel * max_j |Gref[i, j]|).

    row_rel is used only against the FAITHFUL back-end: QUADPACK's abserr under-estimates the error of
    the 61-point rule on the C^2 cubic-spline integrands (measured against mpmath: up to 4e-10 absolute at
    N = 50 although every call reports success at 1e-12), so there the 1e-9 gate is taken relative to the
    row scale (the diagonal entry, ~0.7-0.9) instead of entry by entry.
    rel = 1e-9 is the north-star tolerance for matrix entries.  abs_floor covers the oracle's own epsabs
    (1e-12 per adaptive QUADPACK call, src/Integration.cpp:31,66) for entries that are structurally tiny;
    Eref, when given, is the oracle's own summed error estimate for the entry."""
    tol = rel * np.abs(Gref) + abs_floor + (0 if Eref is None else Eref)
    tol = tol + row_rel * np.abs(Gref).max(axis=1, keepdims=True)
    bad = np.abs(G - Gref) > tol
    assert not bad.any(), (f"{what}: {bad.sum()} entries out of tolerance; worst |d|="
                           f"{np.abs(G - Gref)[bad].max():.3e} at {np.argwhere(bad)[0]}")
