import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_PARENT = os.path.join(ROOT, "awkward-0.x_b200")
for p in (ROOT, PKG_PARENT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(params=GOLDEN_NAMES)
def golden(request):
    g = load_golden(request.param)
    g["name"] = request.param
    return g


def same(a, b):
    """Bit-exact equality including dtype, NaN positions and the sign of zero."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        if np.array_equal(a.view("u%d" % a.dtype.itemsize), b.view("u%d" % b.dtype.itemsize)):
            return True
        na, nb = np.isnan(a), np.isnan(b)      # NaN payload / sign bits are not part of the contract
        return bool(np.array_equal(na, nb) and np.array_equal(a[~na], b[~nb]) and
                    np.array_equal(np.signbit(a[~na]), np.signbit(b[~nb])))
    return bool(np.array_equal(a, b))
