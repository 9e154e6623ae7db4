"""-m gpu: ufunc chains on the columns of pairs() / distincts() / cross() results are evaluated by ONE fused kernel
(jg_comb_eval) straight from the two collections.  Checked (1) against the oracle: numpy evaluating the same formula
on the oracle's gathered pair columns, (2) bit for bit against this library's own unfused path (deferred columns
gathered, then one eager kernel per ufunc): the operator functors are shared, so fusion must never change a result."""
import zlib

import numpy as np
import pytest

from conftest import same
from oracle import jagged_oracle as O

pytestmark = pytest.mark.gpu

RTOL = {np.dtype(np.float32): 2e-5, np.dtype(np.float64): 1e-12}


@pytest.fixture(scope="module")
def ak():
    import awkward0_b200
    return awkward0_b200


def collection(ak, counts, dtype, seed):
    rng = np.random.default_rng(seed)
    m = int(counts.sum())
    cols = dict(pt=rng.exponential(20.0, m).astype(dtype), eta=rng.uniform(-2.4, 2.4, m).astype(dtype),
                phi=rng.uniform(-np.pi, np.pi, m).astype(dtype))
    first = ak.JaggedArray.fromcounts(counts, cols["pt"])
    mk = lambda x: ak.JaggedArray.fromoffsets(first.offsets, x)
    return ak.JaggedArray.zip(pt=first, eta=mk(cols["eta"]), phi=mk(cols["phi"])), cols


def mass(one, two):
    return np.sqrt(2 * one.pt * two.pt * (np.cosh(one.eta - two.eta) - np.cos(one.phi - two.phi)))


def host_mass(l, r):
    with np.errstate(all="ignore"):
        return np.sqrt(2 * l["pt"] * r["pt"] * (np.cosh(l["eta"] - r["eta"]) - np.cos(l["phi"] - r["phi"])))


def oracle_sides(kind, off_a, cols_a, off_b, cols_b):
    if kind == "pairs":
        poff, i, j = O.argpairs(off_a)
    elif kind == "distincts":
        poff, i, j = O.argdistincts(off_a)
    else:
        poff, i, j = O.argcross(off_a, off_b)
    left = {k: O.gather_columns(off_a, v, poff, i)[0] for k, v in cols_a.items()}
    right = {k: O.gather_columns(off_b if kind == "cross" else off_a, v, poff, j)[0]
             for k, v in (cols_b if kind == "cross" else cols_a).items()}
    return poff, left, right


def make_counts(kind, n, rng):
    if kind == "poisson2":
        return rng.poisson(2.0, n).astype(np.int64)
    if kind == "poisson5":
        return rng.poisson(5.0, n).astype(np.int64)
    c = rng.poisson(1.5, n).astype(np.int64)               # "long": a few events with hundreds of items
    c[rng.integers(0, n, 6)] = rng.integers(60, 400, 6)
    return c


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kind", ["pairs", "distincts", "cross"])
@pytest.mark.parametrize("shape", ["poisson2", "poisson5", "long"])
def test_fused_mass_matches_oracle_and_unfused(ak, kind, shape, dtype):
    from awkward0_b200 import lazy
    rng = np.random.default_rng(zlib.crc32((kind + shape).encode()))
    n = 3000 if shape == "long" else 40000
    ca, cb = make_counts(shape, n, rng), make_counts("poisson2", n, rng)
    mu, cols_a = collection(ak, ca, dtype, 1)
    el, cols_b = collection(ak, cb, dtype, 2)
    off_a, off_b = O.counts2offsets(ca), O.counts2offsets(cb)

    def run():
        p = mu.cross(el) if kind == "cross" else getattr(mu, kind)()
        return p, mass(p.i0, p.i1)

    p, m = run()
    assert isinstance(m.content, lazy.LazyArray) and m.content.deferred          # nothing has run yet
    assert all(x.deferred for x in (p.i0.pt.content, p.i1.phi.content))
    got = m.content.get()
    assert all(x.deferred for x in (p.i0.pt.content, p.i1.phi.content))           # the gathers were never needed
    poff, left, right = oracle_sides(kind, off_a, cols_a, off_b, cols_b)
    assert np.array_equal(m.offsets.get(), poff)
    want = host_mass(left, right)
    assert got.dtype == want.dtype and got.shape == want.shape
    # cosh - cos cancels for nearly collinear pairs: compare with a float64 evaluation plus an absolute floor
    ref = host_mass({k: v.astype(np.float64) for k, v in left.items()}, {k: v.astype(np.float64) for k, v in right.items()})
    nanmask = np.isnan(ref)
    tol = RTOL[np.dtype(dtype)]
    eps = np.finfo(dtype).eps
    floor = np.sqrt(2 * left["pt"].astype(np.float64) * right["pt"].astype(np.float64) * 8 * eps * np.cosh(2 * 2.4))
    assert np.all(np.abs(got[~nanmask].astype(np.float64) - ref[~nanmask]) <= tol * np.abs(ref[~nanmask]) + floor[~nanmask])
    # the gathered columns themselves are bit-exact
    assert same(p.i0.pt.content.get(), left["pt"]) and same(p.i1.eta.content.get(), right["eta"])

    lazy.ENABLED[0] = False
    try:
        _, m2 = run()
        assert not isinstance(m2.content, lazy.LazyArray)
        unfused = m2.content.get()
    finally:
        lazy.ENABLED[0] = True
    assert same(got, unfused)


def test_fused_expression_zoo(ak):
    """Every fusable operator once, comparisons at the root, constants, a P-sized materialised operand."""
    from awkward0_b200 import lazy
    rng = np.random.default_rng(5)
    counts = rng.poisson(3.0, 20000).astype(np.int64)
    x = rng.uniform(0.1, 2.0, int(counts.sum()))
    a = ak.JaggedArray.fromcounts(counts, x)
    off = O.counts2offsets(counts)
    poff, l, r = O.pairs(off, x)

    def exprs(one, two, npy):
        yield (one - two) ** 2 + npy.abs(one) / (two + 1.5)
        yield npy.arctan2(one, two) + npy.hypot(one, two) - npy.maximum(one, 1.0) * npy.minimum(two, 1.0)
        yield npy.exp(-one) + npy.log(two) + npy.log10(one) + npy.log2(two) + npy.log1p(one) + npy.expm1(two) + npy.exp2(one)
        yield npy.sin(one) * npy.cos(two) + npy.tan(one * 0.5) + npy.sinh(one) - npy.cosh(two) + npy.tanh(one)
        yield npy.arcsin(one * 0.4) + npy.arccos(two * 0.4) + npy.arctan(one) + npy.arcsinh(two)
        yield npy.floor(one * 3) + npy.ceil(two * 3) + npy.rint(one * 7) + npy.sign(one - two) + npy.square(two) + npy.reciprocal(one)
        yield (one * 5) // (two + 0.25) + (one * 5) % (two + 0.25) + (-one) + npy.fabs(two - one) + npy.sqrt(one / two)
        yield one * two > 1.0
        yield (one + two) <= (one * two) * 2

    p = a.pairs()
    for got, want in zip(exprs(p.i0, p.i1, np), exprs(l, r, np)):
        assert isinstance(got.content, lazy.LazyArray) and got.content.deferred
        g = got.content.get()
        assert g.dtype == want.dtype
        if want.dtype == np.bool_:
            assert np.array_equal(g, want)
        else:
            assert np.allclose(g, want, rtol=1e-12, atol=1e-13, equal_nan=True)
    # a materialised P-sized operand mixes in by output position; so does a plain DeviceArray
    s = p.i0 + p.i1
    s.content.get()
    t = (s * p.i0 - 1.0).content
    assert isinstance(t, lazy.LazyArray) and t.deferred
    assert np.allclose(t.get(), (l + r) * l - 1.0, rtol=1e-14)
    # dtype changes and integer columns run eagerly (same answers)
    u = (p.i0 * np.float32(2)).content                     # float64 column x float32 scalar: stays float64, fused
    assert np.array_equal(u.get(), l * np.float32(2))
    ai = ak.JaggedArray.fromcounts(counts, (x * 100).astype(np.int64))
    pi = ai.pairs()
    v = pi.i0 + pi.i1
    assert not isinstance(v.content, lazy.LazyArray)
    assert np.array_equal(v.content.get(), (l * 100).astype(np.int64) + (r * 100).astype(np.int64))
    w = (p.i0 + pi.i1).content                              # float64 + int64 -> eager float64
    assert np.array_equal(w.get(), l + (r * 100).astype(np.int64))


def test_long_chain_is_cut_not_refused(ak):
    from awkward0_b200 import lazy
    counts = np.array([3, 0, 2, 5, 1], dtype=np.int64)
    x = np.linspace(0.5, 1.5, int(counts.sum()))
    a = ak.JaggedArray.fromcounts(counts, x)
    poff, l, r = O.distincts(O.counts2offsets(counts), x)
    p = a.distincts()
    e, want = p.i0, l
    for i in range(70):
        e = e * 1.0009765625 + p.i1 - i
        want = want * 1.0009765625 + r - i
    assert np.array_equal(e.content.get(), want)
    assert np.array_equal(e.offsets.get(), poff)


def test_deferred_columns_behave_like_arrays(ak):
    from awkward0_b200 import lazy
    a = ak.JaggedArray.fromiter([[1.5, 2.5, 3.5], [], [4.5, 5.5]])
    p = a.pairs()
    c = p.i0.content
    assert isinstance(c, lazy.LazyArray) and c.deferred and len(c) == 9 and c.dtype == np.float64 and c.shape == (9,)
    assert p.counts.tolist() == [6, 0, 3] and c.deferred                   # offsets-only questions do not gather
    assert p.i0.sum().tolist() == [1.5 * 3 + 2.5 * 2 + 3.5, 0.0, 4.5 * 2 + 5.5]
    assert not c.deferred and not p.i1.content.deferred                    # one sweep gathered both sides
    assert p.tolist()[2] == [(4.5, 4.5), (4.5, 5.5), (5.5, 5.5)]
    q = a.pairs()
    big = q[q.i0 + q.i1 >= 6.0]                                            # fused comparison used as a jagged mask
    assert big.tolist() == [[(2.5, 3.5), (3.5, 3.5)], [], [(4.5, 4.5), (4.5, 5.5), (5.5, 5.5)]]
    e = ak.JaggedArray.fromcounts([0, 0], np.array([], dtype=np.float32))
    z = e.pairs()
    assert (z.i0 * z.i1).tolist() == [[], []]
    # nested=True keeps the columns deferred one level down
    n = a.cross(a, nested=True)
    d = np.abs(n.i0 - n.i1)
    assert d.tolist() == [[[0.0, 1.0, 2.0], [1.0, 0.0, 1.0], [2.0, 1.0, 0.0]], [], [[0.0, 1.0], [1.0, 0.0]]]
    assert d.min().tolist() == [[0.0, 0.0, 0.0], [], [0.0, 0.0]]


def test_fused_mass_full_size(ak):
    """BASELINE configs[3] at full size (50 M events): size-independent checks of the fused kernel."""
    import torch
    n = 50_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    counts = torch.poisson(torch.full((n,), 2.0, device="cuda"), generator=g).to(torch.int64)
    m = int(counts.sum())
    pt = torch.empty(m, device="cuda", dtype=torch.float32).exponential_(1 / 20.0, generator=g)
    eta = (torch.rand(m, device="cuda", generator=g) * 4.8 - 2.4)
    phi = (torch.rand(m, device="cuda", generator=g) * 2 - 1) * np.pi
    first = ak.JaggedArray.fromcounts(ak.DeviceArray(counts), ak.DeviceArray(pt))
    mk = lambda t: ak.JaggedArray.fromoffsets(first.offsets, ak.DeviceArray(t))
    mu = ak.JaggedArray.zip(pt=first, eta=mk(eta), phi=mk(phi))
    p = mu.pairs()
    mm = mass(p.i0, p.i1)
    c = counts.double()
    assert len(mm.content) == int((c * (c + 1) / 2).sum().item())
    # a muon paired with itself has zero mass; pairs are symmetric: sum over i<=j of m^2 is what the closed form gives
    self_pairs = (p.i0.pt == p.i1.pt) & (p.i0.eta == p.i1.eta) & (p.i0.phi == p.i1.phi)
    zero = mm[self_pairs]
    assert float(zero.content.max()) == 0.0 and len(zero.content) >= m
    d = mu.distincts()
    md = mass(d.i0, d.i1)
    assert len(md.content) == len(mm.content) - m
    assert bool((md.content >= 0).all())
    # the same sum two ways: fused on pairs (self pairs contribute 0) and fused on distincts
    s1, s2 = float((mm * mm).sum().sum()), float((md * md).sum().sum())
    assert abs(s1 - s2) <= 1e-4 * abs(s2)


def test_wide_expression_falls_back_stepwise(ak):
    """Ten distinct columns of one side: the program does not fit (8 per side); the value is still right."""
    rng = np.random.default_rng(8)
    counts = rng.poisson(2.0, 5000).astype(np.int64)
    m = int(counts.sum())
    cols = [rng.normal(size=m).astype(np.float32) for _ in range(10)]
    first = ak.JaggedArray.fromcounts(counts, cols[0])
    rec = ak.JaggedArray.zip(**dict(("c%d" % i, ak.JaggedArray.fromoffsets(first.offsets, c) if i else first) for i, c in enumerate(cols)))
    p = rec.pairs()
    e = p.i0.c0
    for i in range(1, 10):
        e = e + p.i0["c%d" % i]
    poff, left, _ = O.pairs(O.counts2offsets(counts), np.arange(m))
    want = cols[0][left]
    for c in cols[1:]:
        want = want + c[left]
    assert same(e.content.get(), want)


# ---- a[a > c]: the comparison is evaluated inside the selection sweeps (jg_compare_select) -------------------------
def _counts_for(kind, n, rng):
    if kind == "poisson5":
        return rng.poisson(5.0, n).astype(np.int64)
    c = rng.geometric(0.3, n).astype(np.int64) - 1                 # heavy tail: tiles that do not fit the stage
    c[rng.integers(0, n, 4)] = rng.integers(3000, 30000, 4)
    return c


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", ["poisson5", "heavy"])
def test_compare_select_matches_oracle_and_mask_path(ak, shape, dtype):
    from awkward0_b200 import lazy
    rng = np.random.default_rng(31)
    counts = _counts_for(shape, 60000, rng)
    m = int(counts.sum())
    x = rng.normal(20.0, 15.0, m).astype(dtype)
    x[rng.integers(0, m, 200)] = np.nan
    x[rng.integers(0, m, 50)] = np.inf
    x[rng.integers(0, m, 50)] = -np.inf
    x[rng.integers(0, m, 300)] = 20.0                              # ties with the threshold
    off = O.counts2offsets(counts)
    a = ak.JaggedArray.fromcounts(counts, x)
    cases = [(lambda v: v > 20, np.greater), (lambda v: v >= 20, np.greater_equal), (lambda v: v < 20, np.less),
             (lambda v: v <= 20, np.less_equal), (lambda v: 20 < v, np.greater), (lambda v: 20.0 >= v, np.less_equal),
             (lambda v: v > -0.0, np.greater), (lambda v: v > np.inf, np.greater)]
    for f, uf in cases:
        mask = f(a)
        assert isinstance(mask.content, lazy.LazyCompare) and mask.content.deferred
        sel = a[mask]
        assert mask.content.deferred                                # the mask was never written
        with np.errstate(invalid="ignore"):
            hmask = f(x)
        soff, want = O.mask_select(off, x, off, hmask)
        assert np.array_equal(sel.offsets.get(), soff) and same(sel.content.get(), want)
        assert np.array_equal(sel.counts.get(), np.diff(soff))
        assert same(mask.content.get(), hmask)                     # and it still is an ordinary mask when asked
        lazy.ENABLED[0] = False
        try:
            ref = a[f(a)]
        finally:
            lazy.ENABLED[0] = True
        assert np.array_equal(ref.offsets.get(), sel.offsets.get()) and same(ref.content.get(), sel.content.get())


def test_compare_select_other_columns_views_and_writes(ak):
    from awkward0_b200 import lazy
    rng = np.random.default_rng(32)
    counts = rng.poisson(4.0, 30000).astype(np.int64)
    m = int(counts.sum())
    off = O.counts2offsets(counts)
    key32 = rng.normal(size=m).astype(np.float32)
    k = ak.JaggedArray.fromcounts(counts, key32)
    for content in (rng.normal(size=m), rng.integers(-9, 9, m).astype(np.int64), rng.integers(0, 200, m).astype(np.uint8),
                    rng.normal(size=m).astype(np.float32), rng.integers(-9, 9, m).astype(np.int16)):
        a = ak.JaggedArray.fromoffsets(k.offsets, content)
        sel = a[k > 0.25]
        soff, want = O.mask_select(off, content, off, key32 > 0.25)
        assert np.array_equal(sel.offsets.get(), soff) and same(sel.content.get(), want)
    key64 = rng.normal(size=m)
    a64 = ak.JaggedArray.fromoffsets(k.offsets, rng.normal(size=m))           # float64 by another float64: mask path
    sel = a64[ak.JaggedArray.fromoffsets(k.offsets, key64) <= 0.5]
    soff, want = O.mask_select(off, a64.content.get(), off, key64 <= 0.5)
    assert np.array_equal(sel.offsets.get(), soff) and same(sel.content.get(), want)
    # a view whose content starts at offsets[0] != 0
    v = k[1000:]
    sel = v[v > 0.0]
    voff = off[1000:] - off[1000]
    vx = key32[off[1000]:]
    soff, want = O.mask_select(voff, vx, voff, vx > 0.0)
    assert np.array_equal(sel.offsets.get(), soff) and same(sel.content.get(), want)
    # records: every column through the same (materialised once) mask
    rec = ak.JaggedArray.zip(x=k, y=ak.JaggedArray.fromoffsets(k.offsets, key64))
    r = rec[rec.x < 0.0]
    soff, wx = O.mask_select(off, key32, off, key32 < 0.0)
    _, wy = O.mask_select(off, key64, off, key32 < 0.0)
    assert np.array_equal(r.offsets.get(), soff) and same(r.x.content.get(), wx) and same(r.y.content.get(), wy)
    # a deferred comparison sees the data as it was when it was written, even if the array is assigned to later
    b = ak.JaggedArray.fromcounts(counts, key32.copy())
    mask = b > 0.5
    assert mask.content.deferred
    b[b < -0.5] = 9.0                                                           # in-place store
    assert not mask.content.deferred
    assert same(mask.content.get(), key32 > 0.5)
    changed = key32.copy()
    changed[key32 < -0.5] = 9.0
    soff, want = O.mask_select(off, changed, off, key32 > 0.5)
    assert same(b[mask].content.get(), want)
