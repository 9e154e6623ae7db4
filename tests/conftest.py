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


def reduce_close(op, got, want, offsets, content, rtol):
    """north_star's tolerance for float `sum` / `prod` (the reduction order differs from the reference's sequential
    reduceat): |got - want| <= rtol * |want|, rtol = 1e-5 (float32) / 1e-12 (float64).  Events that fail it must be
    explained by the conditioning of the event itself, the same bound the golden test uses:
      sum : |got - want| <= rtol * sum(|x|)        (cancellation between terms of mixed sign)
      prod: relative error <= 2 * n * eps           (n roundings in another order; only beyond n ~ 80 (f32))
    Returns a boolean array (one per event)."""
    got = np.asarray(got)
    want = np.asarray(want)
    with np.errstate(all="ignore"):
        ok = np.isclose(got, want, rtol=rtol, atol=0, equal_nan=True)
        if ok.all() or len(content) == 0:
            return ok
        off = np.asarray(offsets)
        x = np.nan_to_num(np.asarray(content).astype(np.float64), nan=0.0 if op == "sum" else 1.0, posinf=0.0, neginf=0.0)
        starts = np.minimum(off[:-1], len(x) - 1)
        nonempty = off[1:] > off[:-1]
        if op == "sum":
            absum = np.where(nonempty, np.add.reduceat(np.abs(x), starts), 0.0)
            ok |= np.abs(got.astype(np.float64) - want.astype(np.float64)) <= rtol * absum
        else:
            n = (off[1:] - off[:-1]).astype(np.float64)
            bound = np.maximum(rtol, 2.0 * n * np.finfo(want.dtype).eps)
            finite = np.isfinite(want) & np.isfinite(got) & (np.abs(want) > np.finfo(want.dtype).tiny)
            ok |= finite & (np.abs(got.astype(np.float64) - want.astype(np.float64)) <= bound * np.abs(want.astype(np.float64)))
            # a product whose magnitude CAN leave the normal range on the way depends on the order legitimately: it
            # overflows to inf (and inf * 0 = NaN if a zero follows) in one order and not in another, or underflows at a
            # different factor -- also when the final value is representable (factors above 1 first, below 1 later: the
            # reference's sequential order overflows where a tree order does not).  No order's partial product exceeds the
            # product of all factors above 1 (or falls below that of all factors below 1): judge that, in the log
            # domain, on the NON-ZERO factors; an event that holds a zero and cannot overflow is +-0 in every order and
            # must simply match.
            nz = np.abs(x) > 0
            logs = np.where(nz, np.log2(np.where(nz, np.abs(x), 1.0)), 0.0)
            up = np.where(nonempty, np.add.reduceat(np.maximum(logs, 0.0), starts), 0.0)
            down = np.where(nonempty, np.add.reduceat(np.minimum(logs, 0.0), starts), 0.0)
            haszero = np.where(nonempty, np.add.reduceat((~nz).astype(np.int64), starts), 0) > 0
            info = np.finfo(want.dtype)
            ok |= (up > info.maxexp - 2) | (~haszero & (down < info.minexp + 2))
    return ok
