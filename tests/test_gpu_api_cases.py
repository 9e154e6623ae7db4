"""-m gpu: a battery of JaggedArray API expressions, each evaluated by the UNMODIFIED reference in the build
container (oracle/make_golden_api.py -> tests/golden/api_cases.json) and here by awkward0_b200 on the device:
same value (`tolist()`), or the same exception type.  Expressions the GPU class declares outside the hot path
(NotImplementedError, no CPU fallback) must be on the explicit list below."""
import json
import math
import re
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

DATA = json.load(open(os.path.join(GOLDEN_DIR, "api_cases.json")))

# outside the GPU hot path by design: the reference computes these, this class refuses loudly
OUT_OF_SCOPE = set([
    "J([[0]], [[1]], [1, 2]).tolist()",              # multi-dimensional starts / stops (jagged.py:59,163,439,1504-1506)
    "J.fromcounts([[1, 2]], [1, 2, 3])",
])


@pytest.fixture(scope="module")
def env():
    import awkward0_b200 as ak
    scope = {"J": ak.JaggedArray, "np": np, "A": ak}
    for name, code in DATA["setup"]:
        scope[name] = eval(code, scope)
    return scope


def leaf_dtype(v):
    """dtype of an array result, or of the innermost content of a (nested) JaggedArray; None for records / scalars."""
    for _ in range(4):
        name = type(v).__name__
        if name in ("DeviceArray", "LazyArray", "LazyCompare", "DeviceMatrix", "ndarray"):
            return str(v.dtype)
        if name == "JaggedArray":
            v = v.content
        else:
            break
    return None


def listify(v):
    if isinstance(v, tuple):
        return {"tuple": [listify(x) for x in v]}
    out = {"dtype": leaf_dtype(v)}
    if hasattr(v, "tolist"):
        v = v.tolist()
    if isinstance(v, np.generic):
        v = v.item()
    out["value"] = v
    return out


def same(got, want):
    if isinstance(want, str) and want in ("nan", "inf", "-inf"):
        want = float(want)
    if isinstance(want, dict):
        return isinstance(got, dict) and sorted(got) == sorted(want) and all(same(got[k], want[k]) for k in want)
    if isinstance(want, (list, tuple)):
        return isinstance(got, (list, tuple)) and len(got) == len(want) and all(same(g, w) for g, w in zip(got, want))
    if isinstance(want, bool) or isinstance(got, bool):
        return bool(got) == bool(want) and isinstance(got, (bool, np.bool_, int)) and isinstance(want, (bool, int))
    if isinstance(want, float) or isinstance(got, float):
        if isinstance(got, str):
            return False
        got, want = float(got), float(want)
        if math.isnan(want) or math.isnan(got):
            return math.isnan(want) and math.isnan(got)
        if want == 0.0 and got == 0.0:
            return math.copysign(1.0, want) == math.copysign(1.0, got)
        if EXACT[0]:
            return got == want               # min / max / gathers / +, -, *, / ...: correctly rounded, bit for bit
        return math.isclose(got, want, rel_tol=1e-6, abs_tol=0.0)
    return got == want


# Float results are compared EXACTLY unless the expression contains an operation whose result legitimately depends on
# the evaluation order or on the libm in use (north_star: "float sums and products within rtol"): re-associated
# reductions and transcendental functions.  Everything else on this path is data movement or correctly rounded IEEE.
EXACT = [True]
ORDER_OR_LIBM = re.compile(r"sum|prod|mean|var|std|moment|sqrt|exp|log|sin|cos|tan|power|\*\*|hypot|arctan|arcsin|arccos|cbrt")


def evaluate(expr, scope, code=None):
    try:
        if code is not None:
            exec(code, scope)
        with np.errstate(all="ignore"):
            return listify(eval(expr, scope))
    except Exception as err:
        return {"raises": type(err).__name__, "message": str(err)}


def check(expr, expect, got):
    EXACT[0] = ORDER_OR_LIBM.search(expr) is None
    if "raises" in got and got["raises"] == "NotImplementedError" and expect.get("raises") != "NotImplementedError":
        assert expr in OUT_OF_SCOPE, "{0}: refused ({1}) but not declared out of scope".format(expr, got["message"])
        return
    if "raises" in expect:
        assert got.get("raises") == expect["raises"], "{0}: reference raises {1}, got {2}".format(expr, expect["raises"], got)
        if expect.get("message"):            # the wording too (an empty message, e.g. a bare NotImplementedError, is not pinned)
            assert got["message"] == expect["message"], "{0}: message {1!r}, reference {2!r}".format(expr, got["message"], expect["message"])
        return
    assert "raises" not in got, "{0}: raised {1}: {2}".format(expr, got.get("raises"), got.get("message"))
    if "tuple" in expect:
        assert "tuple" in got and len(got["tuple"]) == len(expect["tuple"]), expr
        for g, w in zip(got["tuple"], expect["tuple"]):
            assert same(g["value"], w["value"]), "{0}: {1} != {2}".format(expr, g["value"], w["value"])
            if w.get("dtype") is not None and g.get("dtype") is not None:
                assert g["dtype"] == w["dtype"], "{0}: dtype {1}, reference {2}".format(expr, g["dtype"], w["dtype"])
        return
    assert same(got["value"], expect["value"]), "{0}: got {1}, reference {2}".format(expr, got["value"], expect["value"])
    if expect.get("dtype") is not None and got.get("dtype") is not None:
        assert got["dtype"] == expect["dtype"], "{0}: dtype {1}, reference {2}".format(expr, got["dtype"], expect["dtype"])


@pytest.mark.parametrize("case", DATA["cases"], ids=[c["expr"] for c in DATA["cases"]])
def test_expression(env, case):
    check(case["expr"], case["expect"], evaluate(case["expr"], dict(env)))


@pytest.mark.parametrize("case", DATA["statements"], ids=[c["code"] for c in DATA["statements"]])
def test_statement(env, case):
    check(case["code"], case["expect"], evaluate(case["expr"], dict(env), code=case["code"]))
