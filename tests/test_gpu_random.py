"""-m gpu: differential tests CUDA-vs-oracle on seeded random inputs (sizes the numpy oracle finishes in seconds)
and size-independent properties at BASELINE.json's full sizes (100 M events)."""
import zlib

import numpy as np
import pytest

from conftest import reduce_close, same
from oracle import jagged_oracle as O

pytestmark = pytest.mark.gpu

RTOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


@pytest.fixture(scope="module")
def ak():
    import awkward0_b200
    return awkward0_b200


def make_counts(kind, n, rng):
    if kind == "poisson5":
        return rng.poisson(5.0, n).astype(np.int64)
    if kind == "poisson10":                      # mean counts of 7 - 14: the event tile staged 32 KB at a time (JG_WIDE_STAGE)
        return rng.poisson(10.0, n).astype(np.int64)
    if kind == "poisson20":                      # ... of 14 - 28: 64 KB at a time (JG_WIDER_STAGE)
        return rng.poisson(20.0, n).astype(np.int64)
    if kind == "poisson40":
        return rng.poisson(40.0, n).astype(np.int64)
    if kind == "geometric":                      # heavy tail: mostly short, a few long events
        c = rng.geometric(0.2, n).astype(np.int64) - 1
        c[rng.integers(0, n, 5)] = rng.integers(5000, 70000, 5)
        return c
    if kind == "sparse":                         # mostly empty
        return (rng.random(n) < 0.05).astype(np.int64) * rng.integers(1, 4, n)
    raise ValueError(kind)


CASES = [("poisson5", 300000, np.float32), ("poisson5", 300000, np.float64), ("poisson40", 20000, np.float32),
         ("geometric", 20000, np.float64), ("sparse", 100000, np.float32), ("poisson5", 1000, np.int32),
         ("geometric", 3000, np.int64), ("poisson5", 20000, np.int8), ("geometric", 3000, np.int16),
         ("poisson5", 20000, np.uint16), ("poisson40", 2000, np.uint32), ("geometric", 3000, np.uint64),
         ("poisson5", 20000, np.uint8), ("poisson10", 60000, np.float32), ("poisson10", 30000, np.float64),
         ("poisson10", 30000, np.int32), ("poisson10", 20000, np.uint32), ("poisson10", 20000, np.int16),
         ("poisson20", 40000, np.float32), ("poisson20", 20000, np.int32), ("poisson20", 20000, np.uint32),
         ("poisson20", 20000, np.float64)]


@pytest.mark.parametrize("kind,n,dtype", CASES)
def test_differential(ak, kind, n, dtype):
    rng = np.random.default_rng(zlib.crc32(repr((kind, n, np.dtype(dtype).name)).encode()))      # (hash() of a str changes per process)
    counts = make_counts(kind, n, rng)
    m = int(counts.sum())
    if np.dtype(dtype).kind == "f":
        content = rng.exponential(25.0, m).astype(dtype)
        content[rng.random(m) < 0.01] = np.nan
    else:
        content = rng.integers(-1000, 1000, m).astype(dtype)
    off = O.counts2offsets(counts)
    a = ak.JaggedArray.fromcounts(counts, content)
    assert same(a.offsets.get(), off)
    assert same(a.parents.get(), O.offsets2parents(off))
    assert same(a.localindex.content.get(), O.localindex(off)[1])
    with np.errstate(all="ignore"):
        for op in ("min", "max", "any", "all"):
            assert same(getattr(a, op)().get(), O.reduce(off, content, op)), op
        for op in ("sum", "prod"):
            got, want = getattr(a, op)().get(), O.reduce(off, content, op)
            assert got.dtype == want.dtype
            if want.dtype.kind == "f":
                ok = reduce_close(op, got, want, off, content, RTOL[want.dtype])      # north_star's rtol, not a multiple
                assert ok.all(), (op, got[~ok][:5], want[~ok][:5])
            else:
                assert same(got, want), op
        for ismin in (True, False):
            roff, ridx = O.argminmax(off, content, ismin)
            r = a.argmin() if ismin else a.argmax()
            assert same(r.offsets.get(), roff) and same(r.content.get(), ridx)
        for asc in (False, True):
            roff, ridx = O.argsort(off, content, asc)
            r = a.argsort(asc)
            assert same(r.offsets.get(), roff) and same(r.content.get(), ridx), ("argsort", asc)
    thr = 20 if np.dtype(dtype).kind == "f" else 0
    moff, mask = O.ufunc_broadcast(np.greater, off, content, thr)
    gm = a > thr
    assert same(gm.content.get(), mask)
    soff, sel = O.mask_select(off, content, moff, mask)
    gs = a[gm]
    assert same(gs.offsets.get(), soff) and same(gs.content.get(), sel)
    flat = rng.normal(size=n).astype(dtype if np.dtype(dtype).kind == "f" else np.float64)
    with np.errstate(all="ignore"):
        _, plus = O.ufunc_broadcast(np.add, off, content, flat)
    assert same((a + flat).content.get(), plus)
    assert same(a.tojagged(flat).content.get(), O.tojagged_flat(off, flat))
    lead = a[a.argmax()]
    roff, ridx = O.argminmax(off, content, False)
    goff, gval = O.index_gather(off, content, roff, ridx)
    assert same(lead.offsets.get(), goff) and same(lead.content.get(), gval)


@pytest.mark.parametrize("mean,n", [(2.0, 200000), (5.0, 40000), (12.0, 3000)])
def test_differential_combinatorics(ak, mean, n):
    rng = np.random.default_rng(int(mean * 1000) + n)
    ca, cb = rng.poisson(mean, n).astype(np.int64), rng.poisson(2.0, n).astype(np.int64)
    xa = rng.exponential(20.0, int(ca.sum())).astype(np.float32)
    xb = rng.integers(0, 1000, int(cb.sum())).astype(np.int64)
    oa, ob = O.counts2offsets(ca), O.counts2offsets(cb)
    a, b = ak.JaggedArray.fromcounts(ca, xa), ak.JaggedArray.fromcounts(cb, xb)
    for name, fn in (("argpairs", O.argpairs), ("argdistincts", O.argdistincts)):
        want = fn(oa)
        got = getattr(a, name)()
        assert same(got.offsets.get(), want[0]) and same(got.i0.content.get(), want[1]) and same(got.i1.content.get(), want[2])
    for name, fn in (("pairs", O.pairs), ("distincts", O.distincts)):
        want = fn(oa, xa)
        got = getattr(a, name)()
        assert same(got.offsets.get(), want[0]) and same(got.i0.content.get(), want[1]) and same(got.i1.content.get(), want[2])
    for k in (2, 3):
        want = O.argchoose(oa, k)
        got = a.argchoose(k)
        assert same(got.offsets.get(), want[0])
        for c in range(k):
            assert same(got[str(c)].content.get(), want[1 + c])
    want = O.argcross(oa, ob)
    got = a.argcross(b)
    assert same(got.offsets.get(), want[0]) and same(got.i0.content.get(), want[1]) and same(got.i1.content.get(), want[2])
    want = O.cross(oa, xa, ob, xb)
    got = a.cross(b)
    assert same(got.i0.content.get(), want[1]) and same(got.i1.content.get(), want[2])


def test_edge_layouts(ak):
    J = ak.JaggedArray
    e = J.fromcounts(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float32))
    assert len(e) == 0 and e.sum().tolist() == [] and e.argmax().tolist() == [] and e[e > 1].tolist() == []
    assert e.pairs().tolist() == [] and e.counts.tolist() == [] and e.offsets.tolist() == [0]
    z = J.fromcounts(np.zeros(1000, dtype=np.int64), np.zeros(0, dtype=np.float64))
    assert z.sum().tolist() == [0.0] * 1000 and z.max().tolist() == [-np.inf] * 1000
    assert z.argmax().counts.tolist() == [0] * 1000 and z[z > 0].offsets.tolist() == [0] * 1001
    big = J.fromcounts([0, 1 << 24, 3], np.ones((1 << 24) + 3, dtype=np.float32))     # one huge event
    assert big.sum().tolist() == [0.0, float(1 << 24), 3.0]
    assert big.argmax().tolist() == [[], [0], [0]] and big.counts.tolist() == [0, 1 << 24, 3]
    assert big[big > 0].counts.tolist() == [0, 1 << 24, 3]
    assert int(big.parents[(1 << 24) - 1]) == 1 and int(big.parents[(1 << 24)]) == 2
    nan = J.fromcounts([2, 1], np.array([np.nan, np.nan, 1.0], dtype=np.float32))
    assert nan.argmax().tolist() == [[], [0]] and nan.sum().tolist() == [0.0, 1.0] and nan.max().tolist() == [-np.inf, 1.0]


def test_properties_at_full_size(ak):
    """BASELINE.json configs[1]/[2] size: 100 M events.  No oracle here -- only size-independent invariants."""
    n = 100_000_000
    rng = np.random.default_rng(12345)
    counts = rng.poisson(5.0, n).astype(np.int64)
    m = int(counts.sum())
    pt = rng.standard_exponential(m, dtype=np.float32) * np.float32(25.0)
    a = ak.JaggedArray.fromcounts(counts, pt)
    del counts
    off = a.offsets
    assert int(off[0]) == 0 and int(off[-1]) == m and len(off) == n + 1
    c = a.counts
    assert int(c.sum()) == m and bool((c >= 0).all())
    assert bool((ak.JaggedArray.counts2offsets(c) == off).all())                     # scan(diff(offsets)) round trip
    s = a.sum()
    assert np.isclose(float(s.sum()), float(a.content.sum()), rtol=1e-3)
    assert float(a.max().max()) == float(a.content.max())
    sel = a[a > 20.0]
    k = len(sel.content)
    assert k == int(sel.counts.sum()) == int(sel.offsets[-1])
    assert float(sel.content.min()) > 20.0
    assert np.isclose(k / m, np.exp(-0.8), atol=1e-3)
    again = sel[sel > 20.0]                                                            # idempotence
    assert bool((again.offsets == sel.offsets).all()) and bool((again.content == sel.content).all())
    del again
    lead = a[a.argmax()]
    nonempty = a.counts > 0
    assert bool((lead.counts == nonempty).all())
    mx = a.max()
    picked = a.tojagged(mx)
    assert bool(((a == picked).sum() >= nonempty).all())        # every non-empty event holds its own maximum
    assert int((a == picked).any().sum()) == int(nonempty.sum())
    del picked, lead
    par = a.parents
    assert int(par[0]) >= 0 and int(par[-1]) <= n - 1 and len(par) == m
    assert bool((par[1:] >= par[:-1]).all())                                           # sortedness
    del par
    pairs_total = int(((c * (c + 1)) // 2).sum())
    sub = ak.JaggedArray.fromoffsets(off[:10_000_001], a.content)
    ap = sub.argpairs()
    assert int(ap.offsets[-1]) == int(((sub.counts * (sub.counts + 1)) // 2).sum()) <= pairs_total
    assert bool((ap.i0.content <= ap.i1.content).all()) and int(ap.i0.content.min()) == 0


@pytest.mark.parametrize("shift", [1, 2, 3, 5])
def test_unaligned_views(ak, shift):
    """Device buffers that start at odd element offsets (views): exercises the unaligned head/tail of the TMA
    staging, the scalar fall-backs of the vector kernels and unaligned scan inputs/outputs."""
    rng = np.random.default_rng(100 + shift)
    n = 20011
    counts = rng.poisson(5.0, n).astype(np.int64)
    m = int(counts.sum())
    for dtype in (np.float32, np.float64, np.int32):
        vals = (rng.exponential(25.0, m) if np.dtype(dtype).kind == "f" else rng.integers(-99, 99, m)).astype(dtype)
        big_content = ak.asdevice(np.concatenate([np.zeros(shift, dtype=dtype), vals, np.zeros(7, dtype=dtype)]))
        big_counts = ak.asdevice(np.concatenate([np.zeros(shift, dtype=np.int64), counts]))
        content = big_content[shift:shift + m]
        cview = big_counts[shift:]
        off = O.counts2offsets(counts)
        a = ak.JaggedArray.fromcounts(cview, content)
        assert same(a.offsets.get(), off)
        with np.errstate(all="ignore"):
            assert same(a.max().get(), O.reduce(off, vals, "max"))
            want = O.reduce(off, vals, "sum")
            got = a.sum().get()
            assert same(got, want) if want.dtype.kind != "f" else np.allclose(got, want, rtol=1e-5)
            roff, ridx = O.argminmax(off, vals, False)
            r = a.argmax()
            assert same(r.offsets.get(), roff) and same(r.content.get(), ridx)
        thr = 20 if np.dtype(dtype).kind == "f" else 0
        moff, mask = O.ufunc_broadcast(np.greater, off, vals, thr)
        big_mask = ak.asdevice(np.concatenate([np.ones(shift, dtype=bool), mask]))
        mj = ak.JaggedArray.fromoffsets(a.offsets, big_mask[shift:])
        soff, sel = O.mask_select(off, vals, moff, mask)
        gs = a[mj]
        assert same(gs.offsets.get(), soff) and same(gs.content.get(), sel)
        assert same((a > thr).content.get(), mask)
        # an offsets view that is 8- but not 16-byte aligned
        big_off = ak.asdevice(np.concatenate([np.zeros(1, dtype=np.int64), off]))
        b = ak.JaggedArray.fromoffsets(big_off[1:], content)
        assert same(b.counts.get(), counts) and same(b.parents.get(), O.offsets2parents(off))
        with np.errstate(all="ignore"):
            assert same(b.min().get(), O.reduce(off, vals, "min"))
        assert same(b.argpairs().i1.content.get(), O.argpairs(off)[2])


@pytest.mark.parametrize("n", [1, 511, 512, 513, 3071, 3072, 3073, 6145, 100003])
@pytest.mark.parametrize("cdtype", [np.int64, np.int32])
def test_argmax_with_tile_counts_from_the_scan(ak, n, cdtype):
    """fromcounts' scan also counts the non-empty events per 512-event tile and argmin / argmax use that instead of
    their own pass over the offsets: same answers as the oracle and as an array built by fromoffsets (own pass)."""
    rng = np.random.default_rng(n)
    counts = (rng.poisson(1.2, n) * (rng.random(n) < 0.7)).astype(cdtype)          # many empty events
    m = int(counts.sum())
    x = rng.normal(size=m).astype(np.float32)
    off = O.counts2offsets(counts)
    a = ak.JaggedArray.fromcounts(counts, x)
    assert (a._tile_nonempty is not None) == (n > 0)
    assert np.array_equal(a._tile_nonempty.cpu().numpy(), np.add.reduceat((counts > 0).astype(np.int64), np.arange(0, n, 512)))
    b = ak.JaggedArray.fromoffsets(off, x)
    assert b._tile_nonempty is None
    for ismin in (False, True):
        woff, widx = O.argminmax(off, x, ismin)
        for arr in (a, b):
            r = arr.argmin() if ismin else arr.argmax()
            assert np.array_equal(r.offsets.get(), woff) and np.array_equal(r.content.get(), widx)
    assert same(a[a.argmax()].content.get(), x[(off[:-1][counts > 0] + O.argminmax(off, x, False)[1])])


@pytest.mark.parametrize("k", [2, 3, 4, 5])
@pytest.mark.parametrize("kind,n,dtype", [("poisson5", 100003, np.float32), ("geometric", 5000, np.float64),
                                          ("sparse", 40000, np.int16), ("poisson40", 3000, np.uint8)])
def test_concatenate_axis1_differential(ak, k, kind, n, dtype):
    """jagged.py:1651-1733: k inputs merged event by event (2..4: one jg_segment_merge sweep; 5: scatter per input),
    inputs that are views (offsets[0] != 0), one of them with no content at all, against the oracle, bit for bit."""
    rng = np.random.default_rng(4200 + 7 * k + n)
    parts, arrays = [], []
    for q in range(k):
        counts = make_counts(kind, n + q, rng)
        if q == 1:
            counts[:] = 0                                     # an input that contributes nothing
        off = O.counts2offsets(counts)
        vals = (rng.normal(size=off[-1]) * 50).astype(dtype) if np.dtype(dtype).kind == "f" else rng.integers(0, 100, off[-1]).astype(dtype)
        arr = ak.JaggedArray.fromcounts(counts, vals)[q:]     # views: the first q events dropped
        arrays.append(arr)
        parts.append((off[q:], vals))
    want_off, want = O.concatenate_axis1(parts)
    got = arrays[0].concatenate(arrays[1:], axis=1)
    assert same(got.offsets.get(), want_off)
    assert got.content.dtype == want.dtype
    assert same(got.content.get(), want)


def test_more_than_2_to_31_elements(ak):
    """Maximum sizes: 220 M events, 2.2e9 float32 elements (positions beyond int32, byte offsets beyond uint32).
    Inputs are drawn on the device (no 10 GB host copy).  Every operation is run on the WHOLE array and checked
    through size-independent properties, and bit for bit against the oracle on the last events, whose positions
    lie above 2^31 -- once as the tail of the whole-array result, once computed on a view that starts there."""
    import torch
    n, t = 220_000_000, 60_000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(2031)
    counts_t = torch.randint(0, 21, (n,), generator=gen, device="cuda", dtype=torch.int64)
    m = int(counts_t.sum())
    assert m > 2 ** 31 + 10_000_000
    content_t = torch.empty(m, dtype=torch.float32, device="cuda")
    content_t.exponential_(1.0 / 25.0, generator=gen)
    a = ak.JaggedArray.fromcounts(counts_t, content_t)
    off = a.offsets
    assert int(off[-1]) == m and len(off) == n + 1
    off_tail = off.tensor[n - t:].cpu().numpy()                      # absolute offsets of the last t events
    base = int(off_tail[0])
    assert base > 2 ** 31
    host_tail = content_t[base:].cpu().numpy()
    rel = off_tail - base
    tail = a[n - t:]                                                  # a view whose offsets[0] is above 2^31

    s = a.sum()
    assert np.isclose(float(s.tensor.double().sum()), float(content_t.double().sum()), rtol=1e-9)
    want = O.reduce(rel, host_tail, "sum")
    assert np.allclose(s.tensor[n - t:].cpu().numpy(), want, rtol=1e-5, atol=0)
    assert np.allclose(tail.sum().get(), want, rtol=1e-5, atol=0)
    want = O.reduce(rel, host_tail, "max")
    assert same(a.max().tensor[n - t:].cpu().numpy(), want) and same(tail.max().get(), want)
    del s

    roff, ridx = O.argminmax(rel, host_tail, False)
    am = a.argmax()
    nonempty = int((counts_t > 0).sum())
    assert int(am.offsets[-1]) == nonempty == len(am.content)
    assert same(am.content.tensor[nonempty - len(ridx):].cpu().numpy(), ridx)
    ta = tail.argmax()
    assert same(ta.offsets.get(), roff) and same(ta.content.get(), ridx)
    lead = a[am]
    goff, gval = O.index_gather(rel, host_tail, roff, ridx)
    assert same(lead.content.tensor[nonempty - len(gval):].cpu().numpy(), gval)
    assert same(tail[ta].content.get(), gval)
    del am, lead

    sel = a[a > 20.0]
    k = int((content_t > 20.0).sum())
    assert len(sel.content) == k == int(sel.offsets[-1]) and k > 2 ** 29
    moff, mask = O.ufunc_broadcast(np.greater, rel, host_tail, 20.0)
    soff, want = O.mask_select(rel, host_tail, moff, mask)
    assert same(sel.content.tensor[k - len(want):].cpu().numpy(), want)
    assert same(sel.counts.tensor[n - t:].cpu().numpy(), O.offsets2counts(soff))
    ts = tail[tail > 20.0]
    assert same(ts.offsets.get(), soff) and same(ts.content.get(), want)
    bm = ak.JaggedArray.fromoffsets(off, ak.asdevice(content_t > 20.0))       # the mask as an array of its own (M bytes)
    assert same(a[bm].content.tensor[k - len(want):].cpu().numpy(), want)
    del sel, bm

    par = a.parents
    assert len(par) == m and same(par.tensor[base:].cpu().numpy(), O.offsets2parents(rel) + (n - t))
    assert bool((par.tensor[1:] >= par.tensor[:-1]).all())
    del par
    li = a.localindex
    assert same(li.content.tensor[base:].cpu().numpy(), O.localindex(rel)[1])
    del li
    flat_t = ak.asdevice(torch.arange(n, dtype=torch.int64, device="cuda").float())
    flat_tail = np.arange(n - t, n, dtype=np.int64).astype(np.float32)
    tj = a.tojagged(flat_t)
    assert same(tj.content.tensor[base:].cpu().numpy(), O.tojagged_flat(rel, flat_tail))
    plus = a + flat_t
    with np.errstate(all="ignore"):
        _, want = O.ufunc_broadcast(np.add, rel, host_tail, flat_tail)
    assert same(plus.content.tensor[base:].cpu().numpy(), want)
    del tj, plus, flat_t

    poff, pi, pj = O.argpairs(rel)
    tp = tail.argpairs()
    assert same(tp.offsets.get(), poff) and same(tp.i0.content.get(), pi) and same(tp.i1.content.get(), pj)
    _, left, right = O.pairs(rel, host_tail)
    tv = tail.pairs()
    assert same(tv.i0.content.get(), left) and same(tv.i1.content.get(), right)
    srt = tail.argsort()
    assert same(srt.content.get(), O.argsort(rel, host_tail)[1])


def test_one_event_longer_than_2_to_31(ak):
    """A single event of 2^31 + 1000 elements between two empty ones (block-per-event paths with 64-bit positions)."""
    import torch
    m = 2 ** 31 + 1000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(31)
    content_t = torch.empty(m, dtype=torch.float32, device="cuda")
    content_t.exponential_(1.0 / 25.0, generator=gen)
    content_t[m - 7] = 1.0e6                                   # the maximum sits above position 2^31
    content_t[5] = -3.0                                        # the minimum at the front
    a = ak.JaggedArray.fromcounts(np.array([0, m, 0], dtype=np.int64), content_t)
    assert a.max().tolist() == [-np.inf, 1.0e6, -np.inf] and a.min().tolist() == [np.inf, -3.0, np.inf]
    assert a.argmax().tolist() == [[], [m - 7], []] and a.argmin().tolist() == [[], [5], []]
    assert a[a.argmax()].tolist() == [[], [1.0e6], []]
    s = a.sum().get()
    assert s[0] == 0 and s[2] == 0 and np.isclose(s[1], float(content_t.double().sum()), rtol=1e-4)
    k = int((content_t > 20.0).sum())
    sel = a[a > 20.0]                      # an oversize tile: the CTA compacts the run chunk by chunk
    assert sel.counts.tolist() == [0, k, 0] and len(sel.content) == k
    assert bool((sel.content.tensor == content_t[content_t > 20.0]).all())
    del sel
    par = a.parents
    assert int(par.tensor.min()) == 1 and int(par.tensor.max()) == 1
    del par
    li = a.localindex.content.tensor
    assert int(li[0]) == 0 and int(li[-1]) == m - 1 and bool((li[1:] - li[:-1] == 1).all())
    del li
    tj = a.tojagged(np.array([1.5, 2.5, 3.5], dtype=np.float32))
    assert float(tj.content.tensor.min()) == 2.5 == float(tj.content.tensor.max())


# ----------------------------------------------------------------------------------------------------------------
# BASELINE.json configs at (or near) their stated sizes against the oracle, element for element
# ----------------------------------------------------------------------------------------------------------------
def test_c1_full_size_against_the_oracle(ak):
    """configs[0] at its full size: 10 M events, Poisson(5), float32 -- sum within north_star's rtol (conditioning
    bound), max / argmax / a > 20 / a[a > 20] / a[a.argmax()] bit for bit."""
    n = 10_000_000
    rng = np.random.default_rng(12345)
    counts = rng.poisson(5.0, n).astype(np.int64)
    pt = rng.standard_exponential(int(counts.sum()), dtype=np.float32) * np.float32(25.0)
    off = O.counts2offsets(counts)
    a = ak.JaggedArray.fromcounts(counts, pt)
    assert same(a.offsets.get(), off)
    got, want = a.sum().get(), O.reduce(off, pt, "sum")
    ok = reduce_close("sum", got, want, off, pt, RTOL[np.dtype(np.float32)])
    assert got.dtype == want.dtype and ok.all(), (got[~ok][:5], want[~ok][:5])
    assert same(a.max().get(), O.reduce(off, pt, "max"))
    roff, ridx = O.argminmax(off, pt, False)
    r = a.argmax()
    assert same(r.offsets.get(), roff) and same(r.content.get(), ridx)
    goff, gval = O.index_gather(off, pt, roff, ridx)
    lead = a[r]
    assert same(lead.offsets.get(), goff) and same(lead.content.get(), gval)
    moff, mask = O.ufunc_broadcast(np.greater, off, pt, np.float32(20))
    soff, sel = O.mask_select(off, pt, moff, mask)
    for gs in (a[a > 20.0], a[ak.JaggedArray.fromoffsets(off, mask)]):       # the fused comparison and a byte mask
        assert same(gs.offsets.get(), soff) and same(gs.content.get(), sel) and same(gs.counts.get(), O.offsets2counts(soff))


def test_c4_slice_against_the_oracle(ak):
    """configs[3] on a 5 M-event slice: muons Poisson(2) with (pt, eta, phi), electrons Poisson(2): pairs / distincts /
    cross indices and gathered values bit for bit, the fused per-pair invariant mass within rtol 1e-5 of a float64
    evaluation plus the cancellation floor of cosh - cos."""
    n = 5_000_000
    rng = np.random.default_rng(4242)
    ca, cb = rng.poisson(2.0, n).astype(np.int64), rng.poisson(2.0, n).astype(np.int64)
    oa, ob = O.counts2offsets(ca), O.counts2offsets(cb)

    def columns(m):
        return dict(pt=rng.exponential(20.0, m).astype(np.float32), eta=rng.uniform(-2.4, 2.4, m).astype(np.float32),
                    phi=rng.uniform(-np.pi, np.pi, m).astype(np.float32))

    cols_a, cols_b = columns(int(ca.sum())), columns(int(cb.sum()))

    def collection(counts, cols):
        first = ak.JaggedArray.fromcounts(counts, cols["pt"])
        mk = lambda x: ak.JaggedArray.fromoffsets(first.offsets, x)
        return ak.JaggedArray.zip(pt=first, eta=mk(cols["eta"]), phi=mk(cols["phi"]))

    mu, el = collection(ca, cols_a), collection(cb, cols_b)

    def mass(one, two):
        return np.sqrt(2 * one.pt * two.pt * (np.cosh(one.eta - two.eta) - np.cos(one.phi - two.phi)))

    for kind in ("pairs", "distincts", "cross"):
        if kind == "cross":
            poff, i, j = O.argcross(oa, ob)
            arg, val = mu.argcross(el), mu.cross(el)
        else:
            poff, i, j = (O.argpairs if kind == "pairs" else O.argdistincts)(oa)
            arg, val = getattr(mu, "arg" + kind)(), getattr(mu, kind)()
        assert same(arg.offsets.get(), poff) and same(arg.i0.content.get(), i) and same(arg.i1.content.get(), j), kind
        right_off, right_cols = (ob, cols_b) if kind == "cross" else (oa, cols_a)
        left = {k: O.gather_columns(oa, v, poff, i)[0] for k, v in cols_a.items()}
        right = {k: O.gather_columns(right_off, v, poff, j)[0] for k, v in right_cols.items()}
        got = mass(val.i0, val.i1).content.get()                                   # the fused kernel
        assert same(val.i0.pt.content.get(), left["pt"]) and same(val.i1.phi.content.get(), right["phi"]), kind
        l64 = {k: v.astype(np.float64) for k, v in left.items()}
        r64 = {k: v.astype(np.float64) for k, v in right.items()}
        ref = np.sqrt(2 * l64["pt"] * r64["pt"] * (np.cosh(l64["eta"] - r64["eta"]) - np.cos(l64["phi"] - r64["phi"])))
        eps = np.finfo(np.float32).eps
        floor = np.sqrt(2 * l64["pt"] * r64["pt"] * 8 * eps * np.cosh(2 * 2.4))
        assert got.dtype == np.float32 and got.shape == ref.shape
        assert np.all(np.abs(got.astype(np.float64) - ref) <= 1e-5 * np.abs(ref) + floor), kind


# ----------------------------------------------------------------------------------------------------------------
# count distributions with LONG events: the element-tiled variants (csrc/jg_elem.cu) against the oracle
# ----------------------------------------------------------------------------------------------------------------
def long_counts(kind, rng):
    if kind == "poisson40":
        return rng.poisson(40.0, 30000).astype(np.int64)
    if kind == "uniform3000":                     # events of 0 .. 3000 elements: thread, warp and CTA segments mixed
        return rng.integers(0, 3000, 700).astype(np.int64)
    if kind == "huge":                            # a few events of 1e5 .. 3e5 elements between empty ones: many tiles per event
        c = np.zeros(40, dtype=np.int64)
        c[rng.choice(40, 7, replace=False)] = rng.integers(100000, 300000, 7)
        return c
    if kind == "tile_edges":                      # events that end exactly on 4096- / 2048-element tile boundaries, empties there
        return np.array([4096, 0, 0, 2048, 2048, 0, 8192, 1, 4095, 0, 4096 * 3, 0, 0], dtype=np.int64)
    if kind == "one_event":
        return np.array([1 << 20], dtype=np.int64)
    if kind == "sparse_long":                     # 99 % empty events and a few long ones: tiles that own thousands of events
        c = np.zeros(200000, dtype=np.int64)
        c[rng.choice(200000, 2000, replace=False)] = rng.integers(500, 3000, 2000)
        return c
    raise ValueError(kind)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("kind", ["poisson40", "uniform3000", "huge", "tile_edges", "one_event", "sparse_long"])
def test_long_event_variants(ak, kind, dtype):
    import zlib
    rng = np.random.default_rng(zlib.crc32((kind + np.dtype(dtype).name).encode()))
    counts = long_counts(kind, rng)
    m = int(counts.sum())
    assert m > 7 * len(counts)                    # the host picks the element-tiled kernels
    if np.dtype(dtype).kind == "f":
        content = rng.exponential(25.0, m).astype(dtype)
        content[rng.random(m) < 0.001] = np.nan
        content[rng.random(m) < 0.001] = -0.0
    else:
        content = rng.integers(-1000, 1000, m).astype(dtype)
    off = O.counts2offsets(counts)
    for shift in (0, 3):                          # a dense array, and a view that starts at event 3 (offsets[0] != 0)
        if shift >= len(counts):
            continue
        a = ak.JaggedArray.fromcounts(counts, content)
        o = off
        if shift:
            a, o = a[shift:], off[shift:]
        assert same(a.parents.get(), O.offsets2parents(o))
        assert same(a.localindex.content.get(), O.localindex(o)[1])
        with np.errstate(all="ignore"):
            for op in ("min", "max", "any", "all"):
                assert same(getattr(a, op)().get(), O.reduce(o, content, op)), (kind, op)
            for op in ("sum", "prod"):
                got, want = getattr(a, op)().get(), O.reduce(o, content, op)
                assert got.dtype == want.dtype
                if want.dtype.kind == "f":
                    ok = reduce_close(op, got, want, o, content, RTOL[want.dtype])
                    assert ok.all(), (kind, op, got[~ok][:5], want[~ok][:5])
                else:
                    assert same(got, want), (kind, op)
            for ismin in (True, False):
                roff, ridx = O.argminmax(o, content, ismin)
                r = a.argmin() if ismin else a.argmax()
                assert same(r.offsets.get(), roff) and same(r.content.get(), ridx), (kind, ismin)
        thr = 20 if np.dtype(dtype).kind == "f" else 0
        moff, mask = O.ufunc_broadcast(np.greater, o, content, thr)
        soff, sel = O.mask_select(o, content, moff, mask)
        for gs in (a[a > thr], a[ak.JaggedArray.fromoffsets(o - o[0], mask)]):       # fused comparison (floats) and byte mask
            assert same(gs.offsets.get(), soff) and same(gs.content.get(), sel), kind
            assert same(gs.counts.get(), O.offsets2counts(soff)), kind
        flat = rng.normal(size=len(o) - 1).astype(dtype if np.dtype(dtype).kind == "f" else np.float64)
        assert same(a.tojagged(flat).content.get(), O.tojagged_flat(o, flat))
        lead = a[a.argmax()]
        roff, ridx = O.argminmax(o, content, False)
        goff, gval = O.index_gather(o, content, roff, ridx)
        assert same(lead.offsets.get(), goff) and same(lead.content.get(), gval)


# ----------------------------------------------------------------------------------------------------------------
# combinatorics: the dimuon-like kernels (pair table / owner table + walk) and the long-event kernels (direct tables,
# include/jagged_b200.h JG_COMB_LONG) must both give the oracle's result on EVERY distribution -- the hint only
# chooses the kernel
# ----------------------------------------------------------------------------------------------------------------
def comb_counts(kind, rng):
    if kind == "poisson2":
        return rng.poisson(2.0, 60000).astype(np.int64)
    if kind == "poisson6":                        # tiles of "choose 4" hold several owner-table chunks
        return rng.poisson(6.0, 6000).astype(np.int64)
    if kind == "around_the_table":                # events at and beyond the look-up table's limits (10 / 12 on chip, 14 - 24 in memory)
        return rng.integers(8, 30, 300).astype(np.int64)
    if kind == "mixed_long":                      # short events with events of 40 - 400 items among them, empties, a tail of empties
        c = rng.poisson(3.0, 5000).astype(np.int64)
        c[rng.choice(5000, 40, replace=False)] = rng.integers(40, 400, 40)
        c[rng.choice(5000, 500, replace=False)] = 0
        c[-700:] = 0
        return c
    if kind == "one_big":                         # one event whose pairs exceed 2^20 among ordinary ones
        c = rng.poisson(2.0, 1500).astype(np.int64)
        c[700] = 1500
        return c
    raise ValueError(kind)


@pytest.mark.parametrize("variant", ["short", "long"])
@pytest.mark.parametrize("kind", ["poisson2", "poisson6", "around_the_table", "mixed_long", "one_big"])
def test_combinatorics_kernel_variants(ak, kind, variant, monkeypatch):
    from awkward0_b200 import _lib
    monkeypatch.setattr(_lib, "COMB_LONG_MEAN", -1 if variant == "long" else 1 << 60)
    rng = np.random.default_rng(zlib.crc32(kind.encode()))
    ca = comb_counts(kind, rng)
    cb = rng.poisson(2.0, len(ca)).astype(np.int64)
    if kind == "mixed_long":
        cb[rng.choice(len(cb), 5, replace=False)] = rng.integers(4000, 5000, 5)      # right-hand events of 4096 and more items
    xa = rng.exponential(20.0, int(ca.sum())).astype(np.float32)
    xb = rng.integers(0, 1000, int(cb.sum())).astype(np.int64)
    oa, ob = O.counts2offsets(ca), O.counts2offsets(cb)
    a, b = ak.JaggedArray.fromcounts(ca, xa), ak.JaggedArray.fromcounts(cb, xb)
    for name, fn in (("argpairs", O.argpairs), ("argdistincts", O.argdistincts)):
        want = fn(oa)
        got = getattr(a, name)()
        assert same(got.offsets.get(), want[0]) and same(got.i0.content.get(), want[1]) and same(got.i1.content.get(), want[2]), name
    for name, fn in (("pairs", O.pairs), ("distincts", O.distincts)):
        want = fn(oa, xa)
        got = getattr(a, name)()
        assert same(got.offsets.get(), want[0]) and same(got.i0.content.get(), want[1]) and same(got.i1.content.get(), want[2]), name
        # the fused expression on the same combinations, bit for bit what the gathered columns give
        fused = (got.i0 * 2 - got.i1).content.get()
        assert same(fused, want[1] * np.float32(2) - want[2]), name
    small = kind not in ("mixed_long", "one_big")       # (k-combinations of a 400-item event do not fit any memory)
    for k in ((2, 3, 4, 5) if small else (2,)):
        want = O.argchoose(oa, k)
        got = a.argchoose(k)
        assert same(got.offsets.get(), want[0]), k
        for c in range(k):
            assert same(got[str(c)].content.get(), want[1 + c]), (k, c)
    want = O.argcross(oa, ob)
    got = a.argcross(b)
    assert same(got.offsets.get(), want[0]) and same(got.i0.content.get(), want[1]) and same(got.i1.content.get(), want[2])
    want = O.cross(oa, xa, ob, xb)
    got = a.cross(b)
    assert same(got.i0.content.get(), want[1]) and same(got.i1.content.get(), want[2])
    # records: every leaf column in one sweep (jg_comb_gather_multi)
    rec = ak.JaggedArray.zip(x=a, y=ak.JaggedArray.fromoffsets(a.offsets, (xa * 3).astype(np.float32)))
    want = O.pairs(oa, xa)
    got = rec.pairs()
    assert same(got.i0.x.content.get(), want[1]) and same(got.i1.y.content.get(), (want[2] * np.float32(3)).astype(np.float32))


@pytest.mark.parametrize("shift", [0, 1, 3])
@pytest.mark.parametrize("L", [0, 1, 31, 32, 33, 64, 2048, 4064, 4095, 4096, 4097, 7167, 7168, 7169, 8192])
def test_selection_row_edges(ak, L, shift):
    """The compaction walks a staged run in rows of 32 positions padded with a never-selected item
    (csrc/jg_compact.cuh): runs that end on / one short of / one past a row or the stage, in every 16-byte phase,
    for the event-tiled kernels (512 events share the run, a second tile follows so that the base is not zero) and
    the element-tiled ones (three long events); strict, non-strict and mirrored comparisons, byte masks, everything /
    nothing / a random half selected."""
    rng = np.random.default_rng(7000 + 13 * L + shift)
    layouts = []
    c = np.full(512, L // 512, dtype=np.int64)
    c[:L % 512] += 1
    layouts.append(np.concatenate([rng.permutation(c), rng.poisson(5.0, 700).astype(np.int64)]))
    if L >= 24:
        cuts = np.sort(rng.integers(0, L + 1, 2))
        layouts.append(np.array([cuts[0], cuts[1] - cuts[0], L - cuts[1]], dtype=np.int64))
    for counts in layouts:
        m = int(counts.sum())
        off = O.counts2offsets(counts)
        for dtype in (np.float32, np.float64, np.int32, np.int16):
            isf = np.dtype(dtype).kind == "f"
            vals = (rng.normal(20.0, 10.0, m) if isf else rng.integers(-99, 99, m)).astype(dtype)
            if isf and m:
                vals[rng.random(m) < 0.02] = np.nan
                vals[rng.random(m) < 0.02] = 20.0
            big = ak.asdevice(np.concatenate([np.zeros(shift, dtype=dtype), vals, np.zeros(5, dtype=dtype)]))
            a = ak.JaggedArray.fromcounts(counts, big[shift:shift + m])
            checks = []
            if isf:
                with np.errstate(invalid="ignore"):
                    checks += [(a > 20, vals > 20), (a >= 20, vals >= 20), (a < 20, vals < 20), (a <= 20, vals <= 20),
                               (a > -1e30, vals > -1e30), (a > 1e30, vals > 1e30)]
                    got = [a[a > 20], a[a >= 20], a[a < 20], a[a <= 20], a[a > -1e30], a[a > 1e30]]
            else:
                got = []
            for mask in (rng.random(m) < 0.5, np.ones(m, dtype=bool), np.zeros(m, dtype=bool)):
                bigm = ak.asdevice(np.concatenate([np.ones(shift + 1, dtype=bool), mask]))
                mj = ak.JaggedArray.fromoffsets(a.offsets, bigm[shift + 1:])
                checks.append((mj, mask))
                got.append(a[mj])
            for (_, mask), g in zip(checks, got):
                soff, sel = O.mask_select(off, vals, off, mask)
                assert same(g.offsets.get(), soff), (L, shift, dtype)
                assert same(g.content.get(), sel), (L, shift, dtype)
                assert same(g.counts.get(), O.offsets2counts(soff)), (L, shift, dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int8])
@pytest.mark.parametrize("layout", ["first", "last", "adjacent", "exact", "one_over", "only_long", "between_empties"])
def test_tiles_that_do_not_fit_the_stage(ak, layout, dtype):
    """Event tiles (512 events) whose run exceeds the 16 KB stage because of a few long events among short ones (mean
    count below 7, so the event-tiled kernels run): the reductions take such a tile in staged GROUPS of consecutive events
    and reduce an event longer than the stage with the whole CTA, the selection stages its chunks, parents / localindex /
    broadcast build one owner table per chunk.  Long events at the start / end of a tile, next to each other, of exactly
    the stage's capacity and one more, alone among empty events."""
    rng = np.random.default_rng(zlib.crc32(repr((layout, np.dtype(dtype).name)).encode()))
    stage = 16384 // np.dtype(dtype).itemsize if np.dtype(dtype).itemsize >= 4 else 8192
    n = 512 * 200
    counts = rng.poisson(3.0, n).astype(np.int64)
    tiles = rng.choice(200, 8, replace=False) * 512
    for j, t in enumerate(tiles):
        big = int(rng.integers(stage + 1, 3 * stage)) if j % 2 else int(rng.integers(stage // 2, stage))
        if layout == "first":
            counts[t] = big
        elif layout == "last":
            counts[t + 511] = big
        elif layout == "adjacent":
            counts[t + 200] = big
            counts[t + 201] = int(rng.integers(stage + 1, 2 * stage))
            counts[t + 202] = stage // 3
        elif layout == "exact":
            counts[t + 100] = stage
            counts[t + 300] = stage
        elif layout == "one_over":
            counts[t + 100] = stage + 1
            counts[t + 101] = stage - 1
        elif layout == "only_long":
            counts[t:t + 512] = 0
            counts[t + 17] = big
            counts[t + 400] = stage + 5
        elif layout == "between_empties":
            counts[t + 250:t + 262] = 0
            counts[t + 256] = big
    assert counts.sum() <= 7 * n
    m = int(counts.sum())
    if np.dtype(dtype).kind == "f":
        content = rng.normal(0.0, 30.0, m).astype(dtype)
        content[rng.random(m) < 0.01] = np.nan
    else:
        content = rng.integers(-100, 100, m).astype(dtype)
    off = O.counts2offsets(counts)
    a = ak.JaggedArray.fromcounts(counts, content)
    with np.errstate(all="ignore"):
        for op in ("min", "max", "any", "all"):
            assert same(getattr(a, op)().get(), O.reduce(off, content, op)), op
        want = O.reduce(off, content, "sum")
        got = a.sum().get()
        if want.dtype.kind == "f":
            assert reduce_close("sum", got, want, off, content, RTOL[np.dtype(dtype)]).all()
        else:
            assert same(got, want)
        for ismin in (False, True):
            roff, ridx = O.argminmax(off, content, ismin)
            r = a.argmin() if ismin else a.argmax()
            assert same(r.offsets.get(), roff) and same(r.content.get(), ridx)
        lead = a[a.argmax()]
        loff, lval = O.index_gather(off, content, *O.argminmax(off, content, False))
        assert same(lead.offsets.get(), loff) and same(lead.content.get(), lval)
    thr = 10 if np.dtype(dtype).kind == "f" else 0
    with np.errstate(invalid="ignore"):
        mask = content > thr
    soff, sel = O.mask_select(off, content, off, mask)
    for g in (a[a > thr], a[ak.JaggedArray.fromoffsets(a.offsets, ak.asdevice(mask))]):
        assert same(g.offsets.get(), soff) and same(g.content.get(), sel) and same(g.counts.get(), O.offsets2counts(soff))
    assert same(a.parents.get(), O.offsets2parents(off))
    assert same(a.localindex.content.get(), O.localindex(off)[1])
    flat = rng.integers(-5, 5, n).astype(dtype)
    boff, bval = O.ufunc_broadcast(np.add, off, content, flat)
    b = a + flat
    assert same(b.offsets.get(), boff) and same(b.content.get(), bval)
