"""CPU tests of the host side of the fused combination expressions (awkward0_b200/lazy.py): the expression DAG
-> accumulator program compiler is checked by running the emitted program with a numpy emulator of
csrc/jg_comb.cu:eval_program and comparing with numpy evaluating the same expression directly.  No GPU, no kernels."""
import numpy as np
import pytest

import awkward0_b200 as ak            # noqa: F401  (binds the C ABI; needs the built library, not a GPU)
from awkward0_b200 import _ufunc, lazy
from awkward0_b200._lib import OP


class _FakeTensor(object):
    _next = [1 << 20]

    def __init__(self):
        self._p = _FakeTensor._next[0]
        _FakeTensor._next[0] += 1 << 12

    def data_ptr(self):
        return self._p


class _FakeColumn(object):
    """Stands where a device column stands: a dtype, a length and a device address."""

    def __init__(self, values):
        self.values = values
        self.dtype = values.dtype
        self.itemsize = values.dtype.itemsize
        self.tensor = _FakeTensor()

    def __len__(self):
        return len(self.values)


_BY_CODE = {}
for _table in (_ufunc._BINARY, _ufunc._UNARY):
    for _uf, _name in _table.items():
        _BY_CODE.setdefault(OP[_name], _uf)


def emulate(prog, dtype, columns, ia, ib):
    """numpy restatement of eval_program (an accumulator machine): columns maps a device address to host values."""
    slots, acc = {}, None
    for pc in range(prog.nins):
        w = prog.ins[pc]
        op, kind, rev, idx = w & 255, (w >> 8) & 15, (w >> 12) & 1, w >> 16
        if op == lazy.EV_STORE:
            assert idx < prog.nslots
            slots[idx] = acc
            continue
        uf = None if op == lazy.EV_LOAD else _BY_CODE[op]
        if uf is not None and uf.nin == 1:
            with np.errstate(all="ignore"):
                acc = uf(acc).astype(dtype)
            continue
        if kind == lazy.EV_LEFT:
            b = columns[prog.left[idx]][ia]
        elif kind == lazy.EV_RIGHT:
            b = columns[prog.right[idx]][ib]
        elif kind == lazy.EV_POS:
            b = columns[prog.pos[idx]]
        elif kind == lazy.EV_CONST:
            b = np.full(len(ia), prog.consts[idx], dtype=dtype)
        else:
            assert kind == lazy.EV_SLOT and idx < prog.nslots
            b = slots[idx]
        if uf is None:
            assert pc == 0 or acc is not None
            acc = b
        else:
            with np.errstate(all="ignore"):
                acc = (uf(b, acc) if rev else uf(acc, b)).astype(dtype)
    assert (prog.ins[0] & 255) == lazy.EV_LOAD
    return acc != 0 if prog.out_bool else acc


def make_ctx(dtype, m=50, p=400, seed=3):
    rng = np.random.default_rng(seed)
    ctx = lazy.CombContext(0, None, None, 10, None, p)
    ia, ib = rng.integers(0, m, p), rng.integers(0, m, p)
    cols = {}

    def leaf(side):
        c = _FakeColumn(rng.uniform(0.5, 3.0, m).astype(dtype))
        cols[c.tensor.data_ptr()] = c.values
        return ctx.leaf(side, c), c.values[ia if side == "L" else ib]

    return ctx, ia, ib, cols, leaf


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_invariant_mass_program(dtype):
    ctx, ia, ib, cols, leaf = make_ctx(dtype)
    (pt1, vpt1), (pt2, vpt2) = leaf("L"), leaf("R")
    (eta1, veta1), (eta2, veta2) = leaf("L"), leaf("R")
    (phi1, vphi1), (phi2, vphi2) = leaf("L"), leaf("R")
    m = np.sqrt(2 * pt1 * pt2 * (np.cosh(eta1 - eta2) - np.cos(phi1 - phi2)))
    assert isinstance(m, lazy.LazyArray) and m.deferred and m.dtype == np.dtype(dtype) and len(m) == 400
    prog, cdt = lazy._compile(m)
    assert cdt == np.dtype(dtype) and prog.nins == 14 and prog.nslots == 1 and not prog.out_bool
    want = np.sqrt(2 * vpt1 * vpt2 * (np.cosh(veta1 - veta2) - np.cos(vphi1 - vphi2)))
    got = emulate(prog, dtype, cols, ia, ib)
    assert got.dtype == want.dtype and np.array_equal(got, want)


def test_shared_subexpressions_and_comparison_root():
    ctx, ia, ib, cols, leaf = make_ctx(np.float32)
    (x, vx), (y, vy) = leaf("L"), leaf("R")
    d = x - y
    e = d * d + np.abs(d) / (d * d + 1.5)             # d and d*d are shared nodes? (d*d is built twice: two nodes)
    prog, _ = lazy._compile(e)
    stores = [prog.ins[i] & 255 for i in range(prog.nins)].count(lazy.EV_STORE)
    assert stores >= 1 and prog.nslots <= 3            # d = x - y is computed once and parked in a slot
    vd = vx - vy
    assert np.array_equal(emulate(prog, np.float32, cols, ia, ib), vd * vd + np.abs(vd) / (vd * vd + np.float32(1.5)))
    mask = e > 0.75
    assert mask.dtype == np.bool_ and mask.deferred
    prog, _ = lazy._compile(mask)
    assert prog.out_bool == 1
    assert np.array_equal(emulate(prog, np.float32, cols, ia, ib), (vd * vd + np.abs(vd) / (vd * vd + np.float32(1.5))) > 0.75)


def test_long_chain_stays_within_limits():
    ctx, ia, ib, cols, leaf = make_ctx(np.float64)
    (x, vx), (y, vy) = leaf("L"), leaf("R")
    e, want = x, vx
    for i in range(20):
        e = e * 1.0625 + y
        want = want * 1.0625 + vy
        assert isinstance(e, lazy.LazyArray) and e.deferred
        prog, _ = lazy._compile(e)
        assert prog.nins <= lazy.MAX_INS and prog.nslots <= lazy.MAX_SLOTS
        assert np.array_equal(emulate(prog, np.float64, cols, ia, ib), want)


def test_what_is_not_fused():
    ctx, ia, ib, cols, leaf = make_ctx(np.float32)
    (x, _), (y, _) = leaf("L"), leaf("R")
    other = lazy.CombContext(0, None, None, 10, None, 400)
    z = other.leaf("L", _FakeColumn(np.ones(50, np.float32)))
    assert lazy.try_fuse("ADD", [x, z], np.dtype(np.float32), np.dtype(np.float32)) is None      # two contexts
    assert lazy.try_fuse("ADD", [x, 1.0], np.dtype(np.float64), np.dtype(np.float64)) is None     # dtype would change
    assert lazy.try_fuse("AND", [x, y], np.dtype(np.float32), np.dtype(np.float32)) is None       # not a float operator
    lazy.ENABLED[0] = False
    try:
        assert lazy.try_fuse("ADD", [x, y], np.dtype(np.float32), np.dtype(np.float32)) is None
    finally:
        lazy.ENABLED[0] = True
    assert lazy.try_fuse("ADD", [x, y], np.dtype(np.float32), np.dtype(np.float32)) is not None


def test_register_pressure_is_bounded():
    ctx, ia, ib, cols, leaf = make_ctx(np.float32)
    leaves = [leaf("L" if i % 2 else "R") for i in range(8)]
    # a right-leaning sum keeps one running value; a balanced tree of 8 leaves needs 4 live values: both fit
    e = leaves[0][0]
    want = leaves[0][1]
    for l, v in leaves[1:]:
        e = e + l
        want = want + v
    prog, _ = lazy._compile(e)
    assert prog.nslots == 0                           # a chain never leaves the accumulator
    assert np.array_equal(emulate(prog, np.float32, cols, ia, ib), want)
    bal = ((leaves[0][0] * leaves[1][0]) + (leaves[2][0] * leaves[3][0])) - ((leaves[4][0] * leaves[5][0]) + (leaves[6][0] * leaves[7][0]))
    wbal = ((leaves[0][1] * leaves[1][1]) + (leaves[2][1] * leaves[3][1])) - ((leaves[4][1] * leaves[5][1]) + (leaves[6][1] * leaves[7][1]))
    prog, _ = lazy._compile(bal)
    assert prog.nslots <= 3
    assert np.array_equal(emulate(prog, np.float32, cols, ia, ib), wbal)


def test_wide_expression_is_caught_when_needed_not_when_built():
    """More than 8 distinct columns of one side do not fit one program: building stays cheap (no test-compile for
    small trees) and the compiler says so, which makes LazyArray._materialise evaluate the last step eagerly."""
    ctx, ia, ib, cols, leaf = make_ctx(np.float32)
    e = leaf("L")[0]
    for i in range(9):
        e = e + leaf("L")[0]
    assert e.deferred and e._size <= lazy._UNCHECKED_SIZE
    assert lazy._compile(e) is None                      # 10 left columns
    assert lazy._compile(e._args[1]) is None and lazy._compile(e._args[1]._args[1]) is not None      # 9 / 8 columns
