"""Build-container only (skipped where /root/reference is absent, e.g. on the GPU box): randomised cross-checks of
oracle/jagged_oracle.py against the UNMODIFIED reference imported through oracle/_refloader.py.  The committed golden
vectors pin the oracle on fixed inputs; this pins it on fresh seeded ones every time the CPU suite runs here, so the
`-m gpu` differential tests (which compare the CUDA path with the oracle on their own random inputs) rest on an
oracle that has been compared with the reference on the same kinds of inputs."""
import numpy as np
import pytest

from conftest import same
from oracle import _refloader
from oracle import jagged_oracle as O

pytestmark = pytest.mark.skipif(not _refloader.available(), reason="the reference tree only exists in the build container")


@pytest.fixture(scope="module")
def J():
    return _refloader.load().JaggedArray


def make(rng, n, mean, dtype, nan=False):
    counts = rng.poisson(mean, n).astype(np.int64)
    counts[rng.integers(0, n, max(1, n // 50))] = 0
    m = int(counts.sum())
    if np.dtype(dtype).kind == "f":
        content = (rng.normal(size=m) * 30).astype(dtype)
        if nan and m:
            content[rng.random(m) < 0.05] = np.nan
    elif np.dtype(dtype).kind == "b":
        content = rng.random(m) < 0.5
    else:
        content = rng.integers(-50, 50, m).astype(dtype) if np.dtype(dtype).kind == "i" else rng.integers(0, 100, m).astype(dtype)
    return counts, O.counts2offsets(counts), content


CASES = [(7, 300, 5.0, np.float32), (8, 300, 5.0, np.float64), (9, 100, 40.0, np.float32), (10, 400, 0.7, np.int64),
         (11, 200, 3.0, np.int8), (12, 200, 3.0, np.uint16), (13, 150, 4.0, np.bool_), (14, 1, 9.0, np.float64)]


@pytest.mark.parametrize("seed,n,mean,dtype", CASES)
def test_offsets_and_reducers(J, seed, n, mean, dtype):
    rng = np.random.default_rng(seed)
    counts, off, content = make(rng, n, mean, dtype, nan=True)
    a = J.fromcounts(counts, content)
    assert same(O.offsets2parents(off), a.parents)
    assert same(O.localindex(off)[1], a.localindex.content)
    assert same(O.offsets2counts(off), a.counts)
    with np.errstate(all="ignore"):
        for op in ("sum", "prod", "min", "max", "any", "all"):
            assert same(O.reduce(off, content, op), getattr(a, op)()), op
        if np.dtype(dtype).kind != "b":
            for ismin in (True, False):
                roff, ridx = O.argminmax(off, content, ismin)
                r = a.argmin() if ismin else a.argmax()
                assert same(O.offsets2counts(roff), r.counts) and same(ridx, r.flatten())
                goff, gval = O.index_gather(off, content, roff, ridx)
                assert same(gval, a[r].flatten()) and same(O.offsets2counts(goff), a[r].counts)


@pytest.mark.parametrize("seed,n,mean,dtype", CASES[:6])
def test_selection_and_broadcast(J, seed, n, mean, dtype):
    rng = np.random.default_rng(100 + seed)
    counts, off, content = make(rng, n, mean, dtype, nan=True)
    a = J.fromcounts(counts, content)
    thr = 5
    with np.errstate(all="ignore"):
        moff, mask = O.ufunc_broadcast(np.greater, off, content, thr)
        ref_mask = a > thr
        assert same(mask, ref_mask.flatten())
        soff, sel = O.mask_select(off, content, moff, mask)
        assert same(sel, a[ref_mask].flatten()) and same(O.offsets2counts(soff), a[ref_mask].counts)
        flat = rng.normal(size=n).astype(dtype if np.dtype(dtype).kind == "f" else np.float64)
        _, plus = O.ufunc_broadcast(np.add, off, content, flat)
        assert same(plus, (a + flat).flatten())
        assert same(O.tojagged_flat(off, flat), a.tojagged(flat).flatten())
        for asc in (False, True):
            if np.dtype(dtype).kind != "b":
                assert same(O.argsort(off, content, asc)[1], a.argsort(ascending=asc).flatten()), asc
        for length, clip in ((3, False), (2, True)):
            poff, data = O.pad_fillna(off[:-1], off[1:], content, length, 7, clip)
            want = a.pad(length, clip=clip).fillna(7)
            assert same(O.offsets2counts(poff), want.counts) and same(data, np.asarray(want.flatten()))


@pytest.mark.parametrize("seed,n,mean", [(21, 400, 2.0), (22, 120, 5.0), (23, 30, 12.0)])
def test_combinatorics(J, seed, n, mean):
    rng = np.random.default_rng(seed)
    counts, off, content = make(rng, n, mean, np.float64)
    counts_b, off_b, content_b = make(rng, n, mean, np.int64)
    a, b = J.fromcounts(counts, content), J.fromcounts(counts_b, content_b)
    for name, fn in (("argpairs", O.argpairs), ("argdistincts", O.argdistincts)):
        poff, i, j = fn(off)
        ref = getattr(a, name)()
        assert same(O.offsets2counts(poff), ref.counts) and same(i, ref.i0.flatten()) and same(j, ref.i1.flatten()), name
    poff, i, j = O.argcross(off, off_b)
    ref = a.argcross(b)
    assert same(O.offsets2counts(poff), ref.counts) and same(i, ref.i0.flatten()) and same(j, ref.i1.flatten())
    _, left, right = O.cross(off, content, off_b, content_b)
    assert same(left, a.cross(b).i0.flatten()) and same(right, a.cross(b).i1.flatten())
    _, left, right = O.pairs(off, content)
    assert same(left, a.pairs().i0.flatten()) and same(right, a.pairs().i1.flatten())
    for k in (2, 3, 4):
        out = O.argchoose(off, k)
        ref = a.argchoose(k)
        assert same(O.offsets2counts(out[0]), ref.counts)
        for c in range(k):
            assert same(out[1 + c], ref[str(c)].flatten()), (k, c)


@pytest.mark.parametrize("k", [2, 3, 4, 5])
def test_concatenate(J, k):
    rng = np.random.default_rng(300 + k)
    n = 250
    parts, arrays = [], []
    for q in range(k):
        dtype = (np.float64, np.int32, np.float32, np.int64, np.float64)[q]
        counts, off, content = make(rng, n + q, 3.0, dtype)
        if q == 1:
            counts[:] = 0
            off, content = O.counts2offsets(counts), content[:0]
        arrays.append(J.fromcounts(counts, content)[q:])           # views: offsets[0] != 0
        parts.append((off[q:], content))
    off1, cat1 = O.concatenate_axis1(parts)
    ref1 = arrays[0].concatenate(arrays[1:], axis=1)
    assert same(O.offsets2counts(off1), ref1.counts) and same(cat1, ref1.flatten())
    off0, cat0 = O.concatenate_axis0(parts)
    ref0 = arrays[0].concatenate(arrays[1:])
    assert same(O.offsets2counts(off0), ref0.counts) and same(cat0, ref0.flatten())


@pytest.mark.parametrize("seed", [41, 42, 43])
def test_constructors_general_layouts_and_setitem(J, seed):
    rng = np.random.default_rng(seed)
    counts, off, content = make(rng, 200, 3.0, np.float64)
    a = J.fromcounts(counts, content)
    # general layout: reordered, repeated and dropped events over the same content (jagged.py:599-605)
    pick = rng.integers(0, len(counts), 150)
    g = a[pick]
    goff, gcontent = O.compact(off[:-1][pick], off[1:][pick], content)
    assert same(O.offsets2counts(goff), g.counts) and same(gcontent, g.flatten())
    assert same(O.reduce(goff, gcontent, "sum"), g.sum())
    assert O.is_offsets_compatible(off[:-1], off[1:]) and a.iscompact
    # fromparents / fromuniques / fromlocalindex / fromfolding (jagged.py:73-110, 168-242)
    parents = np.sort(rng.integers(-1, 40, 300)).astype(np.int64)
    starts, stops = O.parents2startsstops(parents)
    ref = J.fromparents(parents, np.arange(300))
    assert same(starts, ref.starts) and same(stops, ref.stops)
    uniques = np.sort(rng.integers(0, 30, 200))
    uoff, upar = O.uniques2offsetsparents(uniques)
    rstarts_stops = J.fromuniques(uniques, np.arange(200))
    assert same(uoff[:-1], rstarts_stops.starts) and same(uoff[1:], rstarts_stops.stops) and same(upar, rstarts_stops.parents)
    li = a.localindex.flatten()
    nonempty = counts[counts > 0]
    assert same(O.offsets2counts(O.localindex2offsets(li)), J.fromlocalindex(li, content).counts)
    assert same(O.offsets2counts(O.localindex2offsets(li)), nonempty)
    for length, size in ((17, 5), (20, 5), (3, 7)):
        assert same(O.offsets2counts(O.folding2offsets(length, size)), J.fromfolding(np.arange(length), size).counts)
    # __setitem__ through a jagged mask and through a jagged index (jagged.py:789-838)
    t = J.fromcounts(counts, content.copy())
    m = t > 0
    t[m] = -1.0
    moff, mask = O.ufunc_broadcast(np.greater, off, content, 0)
    assert same(O.setitem(off[:-1], off[1:], content, moff, mask, -1.0), t.content)
    t = J.fromcounts(counts, content.copy())
    idx = J.fromcounts((counts > 0).astype(np.int64), -np.ones(int((counts > 0).sum()), dtype=np.int64))     # last item of every event
    t[idx] = 99.0
    ioff = O.counts2offsets((counts > 0).astype(np.int64))
    assert same(O.setitem(off[:-1], off[1:], content, ioff, idx.content, 99.0), t.content)


@pytest.mark.parametrize("seed,n,mean", [(51, 300, 2.5), (52, 60, 7.0)])
def test_value_combinatorics_and_distincts(J, seed, n, mean):
    rng = np.random.default_rng(seed)
    counts, off, content = make(rng, n, mean, np.float32)
    a = J.fromcounts(counts, content)
    _, left, right = O.distincts(off, content)
    assert same(left, a.distincts().i0.flatten()) and same(right, a.distincts().i1.flatten())
    for k in (2, 3):
        out = O.choose(off, content, k)
        ref = a.choose(k)
        assert same(O.offsets2counts(out[0]), ref.counts)
        for c in range(k):
            assert same(out[1 + c], ref[str(c)].flatten()), (k, c)
