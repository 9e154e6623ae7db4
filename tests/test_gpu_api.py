"""-m gpu: the reference's own test cases (tests/test_jagged.py, tests/test_issues.py) and its exception contract
 run against the CUDA JaggedArray, plus layouts and record content."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ak():
    import awkward0_b200
    return awkward0_b200


CONTENT = [0.0, 1.1, 2.2, 3.3, 4.4, 5.5, 6.6, 7.7, 8.8, 9.9]


def std(ak):
    return ak.JaggedArray([0, 3, 3, 5], [3, 3, 5, 10], CONTENT)


# ---- tests/test_jagged.py -----------------------------------------------------------------------------------
def test_init_and_constructors(ak):        # test_jagged.py:20-38
    J = ak.JaggedArray
    want = [[0.0, 1.1, 2.2], [], [3.3, 4.4], [5.5, 6.6, 7.7, 8.8, 9.9]]
    assert std(ak).tolist() == want
    assert J.fromiter(want).tolist() == want
    assert J.fromoffsets([0, 3, 3, 5, 10], CONTENT).tolist() == want
    assert J.fromcounts([3, 0, 2, 5], CONTENT).tolist() == want
    a = J([], [], [0.0, 1.1, 2.2, 3.3, 4.4])
    assert a[:].tolist() == [] and len(a) == 0
    assert isinstance(std(ak).nbytes, int)


def test_introspection(ak):
    a = ak.JaggedArray.fromcounts([3, 0, 2], np.arange(5, dtype=np.float32))
    assert len(a) == 3 and a.shape == (3,) and a.ndim == 1 and a.dtype == np.dtype(object)
    assert str(a.type) == "[0, 3) -> [0, inf) -> float32"
    assert a.nbytes == 4 * 8 + 5 * 4
    assert a.valid() is True
    assert str(a) == "[[0.0, 1.0, 2.0] [] [3.0, 4.0]]"


def test_tojagged(ak):                     # test_jagged.py:56-64
    J = ak.JaggedArray
    a = J.fromiter([[1], [2, 3], []])
    assert a.tojagged(a).tolist() == a.tolist()
    assert a.tojagged(np.array([3, 4, 5]) + 1).tolist() == [[4], [5, 5], []]
    b = J.fromiter([[], []])
    assert b.tojagged(b).tolist() == b.tolist()
    assert a.tojagged(7).tolist() == [[7], [7, 7], []]


def test_ufunc(ak):                        # test_jagged.py:151-154
    a = std(ak)
    assert (100 + a).tolist() == [[100.0, 101.1, 102.2], [], [103.3, 104.4], [105.5, 106.6, 107.7, 108.8, 109.9]]
    assert (np.array([100, 200, 300, 400]) + a).tolist() == [[100.0, 101.1, 102.2], [], [303.3, 304.4], [405.5, 406.6, 407.7, 408.8, 409.9]]
    assert (a + a).tolist() == [[0.0, 2.2, 4.4], [], [6.6, 8.8], [11.0, 13.2, 15.4, 17.6, 19.8]]
    assert np.sqrt(a).content.get().tolist() == np.sqrt(np.array(CONTENT)).tolist()
    assert (-a).tolist()[2] == [-3.3, -4.4]
    assert ((a > 2) & (a < 8)).tolist() == [[False, False, True], [], [True, True], [True, True, True, False, False]]
    assert (~(a > 2)).tolist()[0] == [True, True, False]


def test_ufunc_dtypes_follow_numpy(ak):
    f32 = ak.JaggedArray.fromcounts([2, 1], np.array([1, 2, 3], dtype=np.float32))
    i32 = ak.JaggedArray.fromcounts([2, 1], np.array([1, 2, 3], dtype=np.int32))
    assert (f32 + 1).content.dtype == np.float32 and (f32 + 1.5).content.dtype == np.float32
    assert (f32 + np.float64(1)).content.dtype == np.float64
    assert (f32 > 20).content.dtype == np.bool_
    assert (i32 + 1).content.dtype == np.int32 and (i32 + 1.5).content.dtype == np.float64
    assert (i32 / i32).content.dtype == np.float64 and (i32 // 2).content.get().tolist() == [0, 1, 1]
    assert (i32 % 2).content.get().tolist() == [1, 0, 1] and (i32 ** 2).content.get().tolist() == [1, 4, 9]
    assert (f32 * np.array([10.0, 100.0], dtype=np.float32)).content.get().tolist() == [10.0, 20.0, 300.0]


def test_pairs_cross_distincts_loops(ak):  # test_jagged.py:214-252
    J = ak.JaggedArray
    for i in (0, 1, 2, 3, 7, 20, 49):
        a = J.fromiter([[], [123], list(range(i)), []])
        c = a.pairs()
        assert c.counts.tolist() == [0, 1, i * (i + 1) // 2, 0]
        left = sum([[x] * (i - x) for x in range(i)], [])
        right = sum([list(range(x, i)) for x in range(i)], [])
        lo, hi = int(c.offsets[2]), int(c.offsets[3])
        assert c.i0.content[lo:hi].tolist() == left and c.i1.content[lo:hi].tolist() == right
        d = a.distincts()
        lo, hi = int(d.offsets[2]), int(d.offsets[3])
        assert d.counts.tolist() == [0, 0, i * (i - 1) // 2, 0]
        assert d.i0.content[lo:hi].tolist() == [x for x, y in zip(left, right) if x != y]
        assert d.i1.content[lo:hi].tolist() == [y for x, y in zip(left, right) if x != y]
    for i in (0, 1, 4, 9):
        for j in (0, 1, 4):
            a = J.fromiter([[], [123], list(range(i)), []])
            b = J.fromiter([[], [456], list(range(j)), [999]])
            c = a.cross(b)
            assert c.counts.tolist() == [0, 1, i * j, 0]
            lo, hi = int(c.offsets[2]), int(c.offsets[3])
            assert c.i0.content[lo:hi].tolist() == np.repeat(range(i), j).tolist()
            assert c.i1.content[lo:hi].tolist() == np.tile(range(j), i).tolist()


def test_argchoose_colex(ak):              # test_jagged.py:254-338
    import itertools
    for k in (2, 3, 4, 5):
        for n in (0, 1, 4, 5, 6, 9, 13):
            a = ak.JaggedArray.fromiter([[], [1234], list(range(n)), []])
            c = a.argchoose(k)
            want = sorted(itertools.combinations(range(n), k), key=lambda t: tuple(reversed(t)))
            assert c.tolist()[2] == want, (k, n)
            assert c.tolist()[0] == [] and c.tolist()[1] == []
    big = ak.JaggedArray.fromcounts([300], np.arange(300))
    got = big.argchoose(3)
    assert len(got.content) == 300 * 299 * 298 // 6
    assert got.tolist()[0][-1] == (297, 298, 299)


def test_pairs_tolist_and_choose_values(ak):
    a = ak.JaggedArray.fromiter([[1.1, 2.2, 3.3], [], [4.4, 5.5]])
    assert a.argdistincts().tolist() == [[(0, 1), (0, 2), (1, 2)], [], [(0, 1)]]
    assert a.distincts().tolist() == [[(1.1, 2.2), (1.1, 3.3), (2.2, 3.3)], [], [(4.4, 5.5)]]
    assert a.choose(2).tolist() == [[(1.1, 2.2), (1.1, 3.3), (2.2, 3.3)], [], [(4.4, 5.5)]]
    assert a.pairs().tolist()[2] == [(4.4, 4.4), (4.4, 5.5), (5.5, 5.5)]


def test_reducers_literals(ak):            # test_jagged.py:384-449
    a = std(ak)
    assert a.sum().tolist() == [3.3000000000000003, 0.0, 7.7, 38.5]
    assert a.prod().tolist() == [0.0, 1.0, 14.52, 24350.911200000002]
    assert a.min().tolist() == [0.0, np.inf, 3.3, 5.5]
    assert a.max().tolist() == [2.2, -np.inf, 4.4, 9.9]
    assert a.argmin().tolist() == [[0], [], [0], [0]]
    assert a.argmax().tolist() == [[2], [], [1], [4]]
    assert a.any().tolist() == [True, False, True, True] and a.all().tolist() == [False, True, True, True]


def test_boolean_indexing(ak):             # test_jagged.py:583-590
    J = ak.JaggedArray
    array1, array2 = J.fromcounts([1], [0]), J.fromcounts([1], [0, 1])
    indices1, indices2 = J.fromcounts([1], [True]), J.fromcounts([1], [True, False])
    assert array1[indices1].tolist() == array1[indices2].tolist() == array2[indices1].tolist() == array2[indices2].tolist() == [[0]]


def test_localindex_parents(ak):           # test_jagged.py:592-602
    a = ak.JaggedArray.fromiter([[1.1, 2.2, 3.3, 4.4, 5.5], [], [6.6, 7.7, 8.8], [9.9]])
    assert a.localindex.tolist() == [[0, 1, 2, 3, 4], [], [0, 1, 2], [0]]
    assert a.parents.tolist() == [0, 0, 0, 0, 0, 2, 2, 2, 3]
    b = ak.JaggedArray([5, 8], [5, 9], a.content)      # a[[False, True, False, True]]
    assert b.localindex.tolist() == [[], [0]]
    assert b.parents.tolist() == [-1, -1, -1, -1, -1, -1, -1, -1, 1]
    assert b.tolist() == [[], [9.9]] and b.sum().tolist() == [0.0, 9.9]


def test_event_level_selection(ak):        # test_jagged.py:592-602 and :99-149 (flat boolean / integer heads)
    a = ak.JaggedArray.fromiter([[1.1, 2.2, 3.3, 4.4, 5.5], [], [6.6, 7.7, 8.8], [9.9]])
    b = a[[False, True, False, True]]
    assert b.tolist() == [[], [9.9]]
    assert b.localindex.tolist() == [[], [0]]
    assert b.parents.tolist() == [-1, -1, -1, -1, -1, -1, -1, -1, 1]
    assert a[a.counts >= 2].tolist() == [[1.1, 2.2, 3.3, 4.4, 5.5], [6.6, 7.7, 8.8]]
    assert a[a.counts >= 2].sum().tolist() == [16.5, 23.1]
    assert a[[2, 1, 0]].tolist() == [[6.6, 7.7, 8.8], [], [1.1, 2.2, 3.3, 4.4, 5.5]]
    assert a[np.array([-1, 0])].tolist() == [[9.9], [1.1, 2.2, 3.3, 4.4, 5.5]]
    assert a[[0, 0, 3]].max().tolist() == [5.5, 5.5, 9.9]
    pairs = a[a.counts >= 2].distincts()
    assert pairs.counts.tolist() == [10, 3]
    with pytest.raises(IndexError):
        a[[4]]
    with pytest.raises(IndexError):
        a[[True, False]]
    c = a.counts
    assert c[c > 0].tolist() == [5, 3, 1] and c[[3, 0]].tolist() == [1, 5]


def test_intra_event_slices_and_moments(ak):     # test_jagged.py:66-97 (a[2:, 1]), :99-149 (a[:, i:j]); base.py:199-248
    lists = [[0.0, 1.1, 2.2], [], [3.3, 4.4], [5.5, 6.6, 7.7, 8.8, 9.9]]
    a = std(ak)
    assert a[2:, 1].tolist() == [4.4, 6.6] and a[2:, -2].tolist() == [3.3, 8.8]
    with pytest.raises(IndexError):
        a[:, 0]
    for start in (None, 0, 1, 2, -1, -2, 7):
        for stop in (None, 0, 1, 3, -1, 9):
            assert a[:, start:stop].tolist() == [x[start:stop] for x in lists], (start, stop)
    assert a[:, :2].sum().tolist() == [1.1, 0.0, 7.7, 12.1]
    assert a[a.counts > 0][:, 0].tolist() == [0.0, 3.3, 5.5]
    b = ak.JaggedArray.fromcounts([3, 0, 2], np.array([1.0, np.nan, 3.0, 0.0, 5.0]))
    assert b.count().tolist() == [2, 0, 2] and b.count_nonzero().tolist() == [3, 0, 1]
    with np.errstate(all="ignore"):
        m = b.mean().tolist()
    assert m[0] == 2.0 and np.isnan(m[1]) and m[2] == 2.5
    c = ak.JaggedArray.fromcounts([2, 3], np.array([1.0, 3.0, 2.0, 4.0, 6.0]))
    assert c.mean().tolist() == [2.0, 4.0]
    assert np.allclose(c.var().tolist(), [1.0, np.var([2.0, 4.0, 6.0])]) and np.allclose(c.std().tolist(), [1.0, np.std([2.0, 4.0, 6.0])])
    assert ak.JaggedArray.fromcounts([2, 1], np.array([3, 0, 5])).count().tolist() == [2, 1]


def test_nested_combinatorics(ak):         # test_jagged.py:340-382
    J = ak.JaggedArray
    a = J.fromiter([[1.1, 2.2, 3.3], [], [4.4, 5.5]])
    b = J.fromiter([[100, 200], [300], [400]])
    assert a.cross(b, nested=True).tolist() == [[[(1.1, 100), (1.1, 200)], [(2.2, 100), (2.2, 200)], [(3.3, 100), (3.3, 200)]], [], [[(4.4, 400)], [(5.5, 400)]]]
    assert a.argcross(b, nested=True).tolist() == [[[(0, 0), (0, 1)], [(1, 0), (1, 1)], [(2, 0), (2, 1)]], [], [[(0, 0)], [(1, 0)]]]
    assert a.pairs(nested=True).tolist() == [[[(1.1, 1.1), (1.1, 2.2), (1.1, 3.3)], [(2.2, 2.2), (2.2, 3.3)], [(3.3, 3.3)]], [], [[(4.4, 4.4), (4.4, 5.5)], [(5.5, 5.5)]]]
    assert a.argpairs(nested=True).tolist() == [[[(0, 0), (0, 1), (0, 2)], [(1, 1), (1, 2)], [(2, 2)]], [], [[(0, 0), (0, 1)], [(1, 1)]]]
    assert a.distincts(nested=True).tolist() == [[[(1.1, 2.2), (1.1, 3.3)], [(2.2, 3.3)]], [], [[(4.4, 5.5)]]]
    assert a.argdistincts(nested=True).tolist() == [[[(0, 1), (0, 2)], [(1, 2)]], [], [[(0, 1)]]]


def test_physics_idiom(ak):                # tests/test_physics.py:18-88 with delta_r written out in ufuncs
    J = ak.JaggedArray

    def particles(pt, eta, phi):
        return J.zip(pt=J.fromiter(pt), eta=J.fromiter(eta), phi=J.fromiter(phi))

    def delta_r(one, two):
        dphi = (one.phi - two.phi + np.pi) % (2 * np.pi) - np.pi
        return np.sqrt((one.eta - two.eta) ** 2 + dphi ** 2)

    jets = particles([[10.0, 20.0, 30.0], [], [40.0, 50.0]], [[-3.0, -2.0, 2.0], [], [-1.0, 1.0]], [[-1.5, 0.0, 1.5], [], [0.78, -0.78]])
    electrons = particles([[20.2, 50.5], [50.5], [50.5]], [[-2.2, 0.0], [0.0], [1.1]], [[0.1, 0.78], [0.78], [-0.77]])
    combinations = jets.cross(electrons, nested=True)
    clean = ~(delta_r(combinations.i0, combinations.i1) < 0.5).any()
    assert clean.tolist() == [[True, False, True], [], [True, False]]
    assert jets[clean].pt.tolist() == [[10.0, 30.0], [], [40.0]]

    gen = particles([[10.0, 20.0, 30.0], [], [40.0, 50.0]], [[-3.0, -2.0, 2.0], [], [-1.0, 1.0]], [[-1.5, 0.0, 1.5], [], [0.78, -0.78]])
    reco = particles([[20.2, 10.1, 30.3, 50.5], [50.5], [50.5]], [[-2.2, -3.3, 2.2, 0.0], [0.0], [1.1]], [[0.1, -1.4, 1.4, 0.78], [0.78], [-0.77]])
    pairing = gen.cross(reco, nested=True)
    metric = delta_r(pairing.i0, pairing.i1)
    index_of_minimized = metric.argmin()
    assert index_of_minimized.tolist() == [[[1], [0], [2]], [], [[0], [0]]]
    passes_cut = metric[index_of_minimized] < 0.5
    assert passes_cut.tolist() == [[[True], [True], [True]], [], [[False], [True]]]
    best = pairing[index_of_minimized][passes_cut]
    genrecos = best.flatten(axis=1)
    assert genrecos.counts.tolist() == [3, 0, 1] and gen.counts.tolist() == [3, 0, 2]
    assert genrecos.i0.pt.tolist() == [[10.0, 20.0, 30.0], [], [50.0]]
    assert genrecos.i1.pt.tolist() == [[10.1, 20.2, 30.3], [], [50.5]]


def test_issues(ak):                       # tests/test_issues.py:16-30
    a = ak.JaggedArray.fromoffsets([2, 4, 4, 7], np.arange(10, dtype=np.float64))
    assert a[a > 3].tolist() == [[], [], [4.0, 5.0, 6.0]]
    assert a[a > 3].sum().tolist() == [0.0, 0.0, 15.0]
    assert a[a > 100][1:].tolist() == [[], []]
    assert a.flatten().tolist() == [2.0, 3.0, 4.0, 5.0, 6.0]
    assert (a + np.array([10.0, 20.0, 30.0])).tolist() == [[12.0, 13.0], [], [34.0, 35.0, 36.0]]
    assert a.pairs().tolist()[0] == [(2.0, 2.0), (2.0, 3.0), (3.0, 3.0)]


def test_noncompact_layout(ak):            # test_jagged.py:499-502 style: starts/stops that are not dense
    content = np.arange(12, dtype=np.float64)
    a = ak.JaggedArray([0, 3, 7, 10], [3, 6, 10, 12], content)     # element 6 is skipped
    assert a.tolist() == [[0.0, 1.0, 2.0], [3.0, 4.0, 5.0], [7.0, 8.0, 9.0], [10.0, 11.0]]
    with pytest.raises(ValueError, match="starts and stops are not compatible with a single offsets array"):
        a.offsets
    assert a.iscompact is False
    assert a.sum().tolist() == [3.0, 12.0, 24.0, 21.0]
    assert a.counts.tolist() == [3, 3, 3, 2]
    assert a.compact().offsets.tolist() == [0, 3, 6, 9, 11]
    assert a.flatten().tolist() == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 7.0, 8.0, 9.0, 10.0, 11.0]
    assert a[a > 4].tolist() == [[], [5.0], [7.0, 8.0, 9.0], [10.0, 11.0]]
    assert a.argmax().tolist() == [[2], [2], [2], [1]]
    rev = ak.JaggedArray([3, 0], [5, 3], content)                   # out-of-order events
    assert rev.tolist() == [[3.0, 4.0], [0.0, 1.0, 2.0]] and rev.max().tolist() == [4.0, 2.0]


def test_index_dtype_policy(ak):
    a = ak.JaggedArray.fromoffsets(np.array([0, 2, 2, 5], dtype=np.int32), np.arange(5, dtype=np.float32))
    assert a.offsets.dtype == np.int32 and a.counts.dtype == np.int32 and a.starts.dtype == np.int32
    assert a.parents.dtype == np.int64 and a.localindex.content.dtype == np.int64
    assert a.argmax().content.dtype == np.int64 and a.pairs().offsets.dtype == np.int64
    assert a[a > 0].offsets.dtype == np.int64
    assert a.sum().tolist() == [1.0, 0.0, 9.0]
    c32 = ak.JaggedArray.fromcounts(np.array([2, 0, 3], dtype=np.int32), np.arange(5))
    assert c32.offsets.dtype == np.int64 and c32.tolist() == [[0, 1], [], [2, 3, 4]]


def test_records_zip_pairs(ak):            # BASELINE configs[3]: dimuon records
    J = ak.JaggedArray
    counts = [2, 0, 3]
    pt = J.fromcounts(counts, np.array([10.0, 20.0, 30.0, 40.0, 50.0], dtype=np.float32))
    eta = J.fromcounts(counts, np.array([0.1, 0.2, 0.3, 0.4, 0.5], dtype=np.float32))
    mu = J.zip(pt=pt, eta=eta)
    assert mu.columns == ["pt", "eta"] and mu.pt.tolist() == pt.tolist()
    p = mu.distincts()
    assert p.counts.tolist() == [1, 0, 3]
    assert p.i0.pt.tolist() == [[10.0], [], [30.0, 30.0, 40.0]] and p.i1.pt.tolist() == [[20.0], [], [40.0, 50.0, 50.0]]
    m = np.sqrt(2 * p.i0.pt * p.i1.pt * (np.cosh(p.i0.eta - p.i1.eta) - 1))
    want = np.sqrt(2 * np.float32(10) * np.float32(20) * (np.cosh(np.float32(0.1) - np.float32(0.2)) - 1))
    assert np.isclose(m.tolist()[0][0], want, rtol=1e-4)
    el = J.zip(pt=J.fromcounts([1, 1, 0], np.array([5.0, 6.0], dtype=np.float32)))
    x = mu.cross(el)
    assert x.counts.tolist() == [2, 0, 0] and x.i0.pt.tolist()[0] == [10.0, 20.0] and x.i1.pt.tolist()[0] == [5.0, 5.0]
    assert mu[mu.pt > 25].pt.tolist() == [[], [], [30.0, 40.0, 50.0]]
    assert mu.sum()["pt"].tolist() == [30.0, 0.0, 120.0]


def test_records_mixed_dtypes(ak):          # record gather groups leaf columns by itemsize (4 / 8 / 1 byte here)
    J = ak.JaggedArray
    counts = [2, 0, 3, 1]
    pt = J.fromcounts(counts, np.arange(6, dtype=np.float32))
    q = J.fromcounts(counts, np.array([1, -1, 1, 1, -1, -1], dtype=np.int64))
    ok = J.fromcounts(counts, np.array([True, False, True, True, False, True]))
    e = J.fromcounts(counts, np.arange(6, dtype=np.float64) * 10)
    mu = J.zip(pt=pt, q=q, ok=ok, e=e)
    p = mu.distincts()
    assert p.i0.pt.tolist() == [[0.0], [], [2.0, 2.0, 3.0], []] and p.i1.pt.tolist() == [[1.0], [], [3.0, 4.0, 4.0], []]
    assert p.i0.q.tolist() == [[1], [], [1, 1, 1], []] and p.i1.q.tolist() == [[-1], [], [1, -1, -1], []]
    assert p.i0.ok.tolist() == [[True], [], [True, True, True], []] and p.i1.e.tolist() == [[10.0], [], [30.0, 40.0, 40.0], []]
    x = mu.cross(J.fromcounts([1, 1, 0, 2], np.array([7, 8, 9, 10], dtype=np.int32)))
    assert x.counts.tolist() == [2, 0, 0, 2]
    assert x.i0.e.tolist() == [[0.0, 10.0], [], [], [50.0, 50.0]] and x.i1.tolist() == [[7, 7], [], [], [9, 10]]
    assert (p.i0.q * p.i1.q).tolist() == [[-1], [], [1, -1, -1], []]


def test_leading_object_fused(ak):           # BASELINE configs[2]: a[a.argmax()] reuses the values of the arg pass
    from awkward0_b200 import _lib
    from oracle import jagged_oracle as orc
    rng = np.random.default_rng(31)
    for dt in (np.float32, np.float64, np.int32, np.int64, np.uint8, np.int16):
        counts = rng.poisson(2.5, 5000)
        counts[rng.integers(0, 5000, 5)] = rng.integers(40, 3000, 5)        # warp- and block-wide events
        content = (rng.normal(0, 50, counts.sum())).astype(dt)
        offsets = orc.counts2offsets(counts)
        a = ak.JaggedArray.fromcounts(counts, content)
        for name, ismin in (("argmax", False), ("argmin", True)):
            arg = getattr(a, name)()
            assert arg._arg_source is not None
            before = _lib.LAUNCHES[0]
            lead = a[arg]
            assert _lib.LAUNCHES[0] == before                                 # no gather pass
            woff, widx = orc.argminmax(offsets, content, ismin)
            assert np.array_equal(lead.offsets.get(), woff)
            want = content[(offsets[:-1][np.diff(woff) > 0]) + widx]
            got = lead.content.get()
            assert got.dtype == want.dtype and np.array_equal(got, want)
            # the same index applied to ANOTHER array of the same shape takes the gather path and agrees
            other = ak.JaggedArray.fromcounts(counts, content.copy())
            before = _lib.LAUNCHES[0]
            lead2 = other[arg]
            assert _lib.LAUNCHES[0] > before
            assert np.array_equal(lead2.content.get(), want) and np.array_equal(lead2.offsets.get(), woff)
            assert lead.counts.tolist() == (counts > 0).astype(int).tolist()
    # NaN events fall back to the exact path (no cached values) and still index correctly
    a = ak.JaggedArray.fromiter([[1.0, np.nan, 3.0], [np.nan], [], [2.0, 5.0]])
    arg = a.argmax()
    assert arg.tolist() == [[2], [], [], [1]] and arg._arg_source is None
    assert a[arg].tolist() == [[3.0], [], [], [5.0]]
    # switch off: plain gather
    ak.JaggedArray.cache_arg_values = False
    try:
        b = ak.JaggedArray.fromcounts([2, 0, 1], [1.0, 4.0, 2.0])
        arg = b.argmax()
        assert arg._arg_source is None and b[arg].tolist() == [[4.0], [], [2.0]]
    finally:
        ak.JaggedArray.cache_arg_values = True


def test_concatenate(ak):                    # test_jagged.py:451-487
    J = ak.JaggedArray
    lst = [[1, 2, 3], [], [4, 5], [6, 7], [8], [], [9], [10, 11], [12]]
    a1, a2, a3 = J.fromiter(lst[:3]), J.fromiter(lst[3:6]), J.fromiter(lst[6:])
    assert a1.concatenate([a2, a3]).tolist() == lst
    assert J.concatenate([a1, a2, a3]).tolist() == lst
    a1 = J([0, 0, 3, 3, 5], [0, 3, 3, 5, 10], [0.0, 1.1, 2.2, 3.3, 4.4, 5.5, 6.6, 7.7, 8.8, 9.9])
    b1 = J([0, 0, 5, 7, 7], [0, 5, 7, 7, 10], [6.5, 7.6, 8.7, 9.8, 10.9, 4.3, 5.4, 1., 2.1, 3.2])
    c1 = a1.concatenate([b1], axis=1)
    assert c1.tolist() == [[], [0., 1.1, 2.2, 6.5, 7.6, 8.7, 9.8, 10.9], [4.3, 5.4], [3.3, 4.4],
                           [5.5, 6.6, 7.7, 8.8, 9.9, 1., 2.1, 3.2]]
    a2, b2 = J([0], [3], [False, False, False]), J([0], [3], [True, True, True])
    assert a2.concatenate([b2], axis=1).content.dtype == np.dtype(bool)
    a3 = a1[[True, True, True, False, True]]
    b3 = b1[[True, True, True, True, False]]
    assert a3.concatenate([b3], axis=1).tolist() == [[], [0., 1.1, 2.2, 6.5, 7.6, 8.7, 9.8, 10.9], [4.3, 5.4],
                                                      [5.5, 6.6, 7.7, 8.8, 9.9]]
    for dt in (np.int32, np.int64, np.float32, np.float64):
        a = J([0], [1], np.ones(1, dtype=dt))
        assert a.concatenate([a], axis=1).content.dtype == np.dtype(dt)
    with pytest.raises(ValueError):
        a1.concatenate([a2], axis=1)
    with pytest.raises(ValueError):
        J.concatenate([])
    # records, general layouts and slices (offsets[0] != 0)
    mu = J.zip(pt=J.fromcounts([2, 0, 1], [1.0, 2.0, 3.0]), q=J.fromcounts([2, 0, 1], [1, -1, 1]))
    both = mu.concatenate([mu[1:]])
    assert both.pt.tolist() == [[1.0, 2.0], [], [3.0], [], [3.0]] and both.q.tolist() == [[1, -1], [], [1], [], [1]]
    side = mu.concatenate([mu], axis=1)
    assert side.pt.tolist() == [[1.0, 2.0, 1.0, 2.0], [], [3.0, 3.0]]
    assert side.q.tolist() == [[1, -1, 1, -1], [], [1, 1]]
    x = J.fromiter(lst)
    assert x[2:7].concatenate([x[5:]], axis=0).tolist() == lst[2:7] + lst[5:]
    assert x[2:6].concatenate([x[5:]], axis=1).tolist() == [p + q for p, q in zip(lst[2:6], lst[5:])]
    # mixed dtypes promote like numpy.concatenate
    f = J.fromcounts([1, 2], np.array([0.5, 1.5, 2.5], dtype=np.float32))
    i = J.fromcounts([2, 0], np.array([7, 8], dtype=np.int64))
    assert J.concatenate([f, i]).content.dtype == np.dtype(np.float64)
    assert J.concatenate([f, i]).tolist() == [[0.5], [1.5, 2.5], [7.0, 8.0], []]
    # nested
    nest = J.fromcounts([2, 1], J.fromcounts([1, 2, 0], [1, 2, 3]))
    assert nest.concatenate([nest]).tolist() == [[[1], [2, 3]], [[]], [[1], [2, 3]], [[]]]


def test_argsort(ak):                        # test_jagged.py:604-607 (masked event left out) + layouts
    J = ak.JaggedArray
    inf, nan = np.inf, np.nan
    a = J.fromiter([[2., 3., 1.], [4., -inf, 5.], [inf, 4., nan, -inf], [nan]])
    assert a.argsort().tolist() == [[1, 0, 2], [2, 0, 1], [0, 1, 3, 2], [0]]
    assert a.argsort(True).tolist() == [[2, 0, 1], [1, 0, 2], [3, 1, 0, 2], [0]]
    assert a[a.argsort()].tolist()[:2] == [[3., 2., 1.], [5., 4., -inf]]
    assert a[1:3].argsort().tolist() == [[2, 0, 1], [0, 1, 3, 2]]
    assert a[[3, 0]].argsort(True).tolist() == [[0], [2, 0, 1]]
    assert J.fromcounts([0, 0], np.zeros(0)).argsort().tolist() == [[], []]
    nest = J.fromcounts([2, 1], J.fromcounts([2, 3, 0], [1, 2, 5, 3, 4]))
    assert nest.argsort().tolist() == [[[1, 0], [0, 2, 1]], [[]]]
    # events around and beyond the shared-memory limits: 2048 / 2049 / 4096 / 4097 / 70001 elements, ties and NaN
    from oracle import jagged_oracle as orc
    rng = np.random.default_rng(77)
    counts = np.array([3, 2048, 0, 2049, 5, 4096, 4097, 1, 70001, 2, 9000], dtype=np.int64)
    for dt in (np.float32, np.float64, np.int16):
        vals = rng.integers(-300, 300, counts.sum()).astype(dt)
        if np.dtype(dt).kind == "f":
            vals[rng.random(len(vals)) < 0.01] = nan
        off = orc.counts2offsets(counts)
        for asc in (False, True):
            got = J.fromcounts(counts, vals).argsort(asc)
            woff, widx = orc.argsort(off, vals, asc)
            assert np.array_equal(got.offsets.get(), woff) and np.array_equal(got.content.get(), widx)


def test_pad_fillna_regular(ak):             # test_jagged.py:189-212, 544-559
    J = ak.JaggedArray
    a = J.fromcounts([3, 0, 2], [1.1, 2.2, 3.3, 4.4, 5.5])
    want = [[1.1, 2.2, 3.3, None], [None, None, None, None], [4.4, 5.5, None, None]]
    assert a.pad(4, clip=True).tolist() == want
    assert a.pad(4).tolist() == want
    assert a.pad(4, clip=True).regular().tolist() == want
    b = J.fromcounts([5, 0, 3, 1], [1.1, 2.2, 3.3, 4.4, 5.5, 6.6, 7.7, 8.8, 9.9])
    assert b.pad(3).tolist() == [[1.1, 2.2, 3.3, 4.4, 5.5], [None, None, None], [6.6, 7.7, 8.8], [9.9, None, None]]
    assert b.pad(3, clip=True).tolist() == [[1.1, 2.2, 3.3], [None, None, None], [6.6, 7.7, 8.8], [9.9, None, None]]
    assert a.pad(4).fillna(999).tolist() == [[1.1, 2.2, 3.3, 999], [999, 999, 999, 999], [4.4, 5.5, 999, 999]]
    assert a.pad(4).fillna(999).regular().tolist() == [[1.1, 2.2, 3.3, 999], [999, 999, 999, 999], [4.4, 5.5, 999, 999]]
    assert a.pad(4, maskedwhen=False).content.maskedwhen is False
    assert a.pad(4, maskedwhen=False).tolist() == want
    r = J.fromcounts([3, 3, 3, 3], np.arange(12.0))
    assert r.regular().tolist() == [[0.0, 1.0, 2.0], [3.0, 4.0, 5.0], [6.0, 7.0, 8.0], [9.0, 10.0, 11.0]]
    assert r.regular().shape == (4, 3) and np.asarray(r.regular()).shape == (4, 3)
    with pytest.raises(ValueError, match="jagged array is not regular"):
        a.regular()
    # general layouts, int content (the value is cast), NaN content (replaced too), empty content
    g = b[[2, 0, 3]]
    assert g.pad(4).fillna(1).tolist() == [[6.6, 7.7, 8.8, 1.0], [1.1, 2.2, 3.3, 4.4, 5.5], [9.9, 1.0, 1.0, 1.0]]
    i = J.fromcounts([1, 2], np.array([5, 6, 7], dtype=np.int32))
    assert i.pad(2).fillna(9.7).tolist() == [[5, 9], [6, 7]] and i.pad(2).fillna(9.7).content.dtype == np.dtype(np.int32)
    n = J.fromcounts([2, 1], np.array([1.0, np.nan, 3.0]))
    assert n.fillna(0).tolist() == [[1.0, 0.0], [3.0]]
    assert n.pad(2).fillna(-1).tolist() == [[1.0, -1.0], [3.0, -1.0]]
    e = J.fromcounts([0, 0], np.zeros(0, dtype=np.float32))
    assert e.pad(2).tolist() == [[None, None], [None, None]] and e.pad(2).fillna(3).tolist() == [[3.0, 3.0], [3.0, 3.0]]
    # records: every column padded with the same mask
    mu = J.zip(pt=J.fromcounts([2, 0, 1], [1.0, 2.0, 3.0]), q=J.fromcounts([2, 0, 1], [1, -1, 1]))
    p = mu.pad(2, clip=True).fillna(0)
    assert p.pt.tolist() == [[1.0, 2.0], [0.0, 0.0], [3.0, 0.0]] and p.q.tolist() == [[1, -1], [0, 0], [1, 0]]
    with pytest.raises(NotImplementedError, match="fillna"):
        a.pad(4).sum()


def test_arrow_ingest(ak):                   # arrow.py:231-430: list offsets + values straight to HBM
    pa = pytest.importorskip("pyarrow")
    rng = np.random.default_rng(9)
    lists = [[1.1, 2.2], [], [3.3], [4.4, 5.5, 6.6]]
    arr = pa.array(lists)
    j = ak.fromarrow(arr)
    assert j.tolist() == lists and j.offsets.dtype == np.dtype(np.int32) and j.content.dtype == np.dtype(np.float64)
    assert (j.sum().get() == np.array([1.1 + 2.2, 0.0, 3.3, 4.4 + 5.5 + 6.6])).all()
    assert ak.fromarrow(arr.slice(1, 3)).tolist() == lists[1:4]                       # sliced: offsets[0] != 0
    big = pa.array([[1, 2], [3], []], type=pa.large_list(pa.int32()))
    jb = ak.fromarrow(big)
    assert jb.tolist() == [[1, 2], [3], []] and jb.offsets.dtype == np.dtype(np.int64) and jb.content.dtype == np.dtype(np.int32)
    # bit-packed booleans at every bit phase
    counts = rng.poisson(3.0, 200)
    flags = [[bool(x) for x in rng.integers(0, 2, c)] for c in counts]
    barr = pa.array(flags, type=pa.list_(pa.bool_()))
    assert ak.fromarrow(barr).tolist() == flags
    for lo in range(1, 12):
        assert ak.fromarrow(barr.slice(lo, 50)).tolist() == flags[lo:lo + 50]
        child = barr.values.slice(lo, 77)
        sub = pa.ListArray.from_arrays(pa.array([0, 7, 7, 77], type=pa.int32()), child)
        assert ak.fromarrow(sub).tolist() == sub.to_pylist()
    # records, nested lists, chunks
    recs = pa.array([[{"pt": 1.5, "q": 1}, {"pt": 2.5, "q": -1}], [], [{"pt": 3.5, "q": 1}]])
    jr = ak.fromarrow(recs)
    assert jr.pt.tolist() == [[1.5, 2.5], [], [3.5]] and jr.q.tolist() == [[1, -1], [], [1]]
    nested = pa.array([[[1, 2], []], [], [[3]]])
    assert ak.fromarrow(nested).tolist() == [[[1, 2], []], [], [[3]]]
    chunked = pa.chunked_array([pa.array(lists[:2]), pa.array([], type=pa.list_(pa.float64())), pa.array(lists[2:])])
    assert ak.fromarrow(chunked).tolist() == lists
    with pytest.raises(NotImplementedError, match="BitMaskedArray"):
        ak.fromarrow(pa.array([[1.0], None]))
    with pytest.raises(NotImplementedError):
        ak.fromarrow(pa.array([1.0, 2.0]))
    # null VALUES: masked content (the reference: BitMaskedArray, arrow.py:262-266); same tolist / fillna as the
    # reference gives for these arrays (checked in the build container)
    for arr, filled in ((pa.array([[1.0, None, 2.0], [], [None]]), [[1.0, 0.0, 2.0], [], [0.0]]),
                        (pa.array([[1, 2, None], [None], [3]], type=pa.list_(pa.int32())), [[1, 2, 0], [0], [3]]),
                        (pa.array([[True, None], [False]]), [[True, False], [False]])):
        jn = ak.fromarrow(arr)
        assert isinstance(jn.content, ak.MaskedArray)
        assert jn.tolist() == arr.to_pylist()
        assert jn.fillna(0).tolist() == filled
        assert jn.counts.tolist() == [len(x) for x in arr.to_pylist()]
        with pytest.raises(NotImplementedError, match="fillna"):
            jn.sum()
    sliced = pa.array([[1.5, None, 2.5], [], [None, 4.5], [6.5]]).slice(1, 3)
    assert ak.fromarrow(sliced).tolist() == [[], [None, 4.5], [6.5]]
    assert ak.fromarrow(sliced).fillna(-1).sum().tolist() == [0.0, 3.5, 6.5]
    # and back
    back = ak.toarrow(j)
    assert back.to_pylist() == lists and back.type == pa.list_(pa.float64())
    assert ak.toarrow(jr).to_pylist() == recs.to_pylist()
    assert ak.toarrow(j[j > 2]).to_pylist() == [[2.2], [], [3.3], [4.4, 5.5, 6.6]]


def test_other_constructors_and_setitem(ak):   # test_jagged.py:33-35, 50-54, 609-657
    J = ak.JaggedArray
    want = [[0.0, 1.1, 2.2], [], [3.3, 4.4], [5.5, 6.6, 7.7, 8.8, 9.9]]
    assert J.fromparents([0, 0, 0, 2, 2, 3, 3, 3, 3, 3], CONTENT).tolist() == want
    u = J.fromuniques([9, 9, 9, 8, 8, 7, 7, 7, 7, 7], CONTENT)
    assert u.tolist() == [[0.0, 1.1, 2.2], [3.3, 4.4], [5.5, 6.6, 7.7, 8.8, 9.9]]
    assert u.parents.tolist() == [0, 0, 0, 1, 1, 2, 2, 2, 2, 2]
    a = J.fromlocalindex([0, 1, 0, 0, 0, 1, 2, 0, 1, 0], CONTENT)
    assert a.tolist() == [[0.0, 1.1], [2.2], [3.3], [4.4, 5.5, 6.6], [7.7, 8.8], [9.9]]
    assert a.starts.tolist() == [0, 2, 3, 4, 7, 9] and a.stops.tolist() == [2, 3, 4, 7, 9, 10]
    assert J.fromlocalindex(a.localindex, CONTENT).tolist() == a.tolist()
    with pytest.raises(ValueError, match="one greater than the previous"):
        J.fromlocalindex([0, 2, 0], [1.0, 2.0, 3.0])
    with pytest.raises(ValueError, match="does not match"):
        J.fromlocalindex(J.fromcounts([1, 2], [0, 0, 1]), [1.0, 2.0, 3.0]).tolist() and J.fromlocalindex(J.fromcounts([2, 1], [0, 0, 0]), [1.0, 2.0, 3.0])
    assert J.fromjagged(std(ak)).tolist() == want
    assert J.fromfolding(np.arange(7), 3).tolist() == [[0, 1, 2], [3, 4, 5], [6]]
    assert J.fromregular(np.arange(12).reshape(4, 3)).tolist() == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 10, 11]]
    assert J.fromregular(np.arange(12).reshape(2, 2, 3)).tolist() == [[[0, 1, 2], [3, 4, 5]], [[6, 7, 8], [9, 10, 11]]]
    for starts, stops, content in (([0, 1, 1], [1, 1, 3], [1.1, 2.2, 3.3]), ([3, 4, 1], [4, 4, 3], [0.0, 2.2, 3.3, 1.1, 4.4])):
        for i1, i2, i3 in (([[True], [], [True, True]], [[False], [], [True, True]], [[False], [], [True, False]]),
                           ([[0], [], [0, 1]], [[], [], [0, 1]], [[], [], [0]])):
            x = J(starts, stops, content)
            x[J.fromiter(i1)] = J.fromiter([[4.4], [], [5.5, 6.6]])
            assert x.tolist() == [[4.4], [], [5.5, 6.6]]
            x[J.fromiter(i2) if any(len(r) for r in i2) else J.fromcounts([0, 0, 0], np.zeros(0, dtype=np.int64))] = [7.7, 8.8]
            assert x.tolist() == [[4.4], [], [7.7, 8.8]]
            x[J.fromiter(i3)] = 9.9
            assert x.tolist() == [[4.4], [], [9.9, 8.8]]
    mu = J.zip(pt=J.fromcounts([2, 0, 1], [1.0, 2.0, 3.0]), q=J.fromcounts([2, 0, 1], [1, -1, 1]))
    mu["w"] = np.array([10.0, 20.0, 30.0])            # one value per event, broadcast (jagged.py:790-791)
    assert mu.w.tolist() == [[10.0, 10.0], [], [30.0]]
    mu["pt2"] = mu.pt * 2
    assert mu.pt2.tolist() == [[2.0, 4.0], [], [6.0]]
    with pytest.raises(IndexError):
        x[J.fromiter([[5], [], []])] = 1.0


def test_inner_slices_with_step(ak):          # jagged.py:634-707: python slice semantics inside every event
    rng = np.random.default_rng(5)
    counts = rng.poisson(4.0, 300)
    counts[:3] = [0, 1, 40]
    rows = [list(map(float, rng.integers(0, 99, c))) for c in counts]
    a = ak.JaggedArray.fromiter(rows)
    for sl in (slice(None, None, 2), slice(None, None, -1), slice(1, None, 2), slice(None, -1, 3), slice(-2, None, -1),
               slice(5, 1, -2), slice(-1, -6, -2), slice(3, None, -3), slice(None, 2, -1), slice(0, 100, 7),
               slice(-100, 100, 2), slice(100, -100, -2), slice(2, 2, 2)):
        assert a[:, sl].tolist() == [r[sl] for r in rows], sl
    b = a[5:200]
    assert b[:, ::-1].tolist() == [r[::-1] for r in rows[5:200]]
    with pytest.raises(ValueError, match="slice step cannot be zero"):
        a[:, ::0]


def test_streamed_ingest(ak):                # ingest.py: host -> HBM with the copy overlapped, event-range views
    import torch
    from oracle import jagged_oracle as orc
    rng = np.random.default_rng(21)
    counts = rng.poisson(5.0, 200000).astype(np.int64)
    counts[1000] = 30000
    pt = rng.exponential(25.0, int(counts.sum()) + 5).astype(np.float32)       # 5 unreachable trailing items
    cp = torch.empty(len(counts), dtype=torch.int64, pin_memory=True)
    cp.numpy()[:] = counts
    pp = torch.empty(len(pt), dtype=torch.float32, pin_memory=True)
    pp.numpy()[:] = pt
    off = orc.counts2offsets(counts)
    want_sum = orc.reduce(off, pt, "sum")
    want_aoff, want_aidx = orc.argminmax(off, pt, False)
    for nchunks, piece in ((1, 1 << 20), (7, 1 << 16), (8, 64 << 20), (13, 4096)):
        st = ak.StreamedJaggedArray(cp, pp, nchunks=nchunks, piece_bytes=piece)
        assert len(st) == len(counts) and st.nchunks == nchunks
        sums, idxs, sel_counts, nsel = [], [], [], 0
        for a in st:
            assert a.counts.dtype == np.dtype(np.int64)
            sums.append(a.sum().get())
            idxs.append(a.argmax().content.get())
            sel = a[a > 20]
            sel_counts.append(sel.counts.get())
            nsel += len(sel.flatten())
        assert np.allclose(np.concatenate(sums), want_sum, rtol=1e-5)
        assert np.array_equal(np.concatenate(idxs), want_aidx)
        moff, mask = orc.ufunc_broadcast(np.greater, off, pt, 20)
        assert nsel == int(mask.sum())
        assert np.array_equal(np.concatenate(sel_counts), orc.reduce_counts_of_mask(off, mask))
        whole = st.whole()
        assert np.array_equal(whole.offsets.get(), off) and np.allclose(whole.sum().get(), want_sum, rtol=1e-5)
    # pageable numpy inputs work too; negative counts raise like fromcounts
    st = ak.StreamedJaggedArray(counts[:1000], pt[:int(counts[:1000].sum())], nchunks=3)
    assert [len(c) for c in st] == [333, 333, 334]
    with pytest.raises(ValueError, match="non-negative"):
        ak.StreamedJaggedArray(np.array([1, -1]), np.zeros(3))
    with pytest.raises(ValueError, match="beyond the length of the content"):
        ak.StreamedJaggedArray(np.array([2, 2]), np.zeros(3))


def test_device_array_surface(ak):
    a = ak.JaggedArray.fromcounts([2, 0, 3], np.arange(5, dtype=np.float32))
    s = a.sum()
    assert isinstance(s, ak.DeviceArray) and s.dtype == np.float32 and len(s) == 3
    assert float(s.sum()) == 10.0 and float(s.max()) == 9.0 and float(s.min()) == 0.0
    assert (s > 0).tolist() == [True, False, True] and bool((s >= 0).all()) and bool((s > 5).any())
    assert (s * 2 + 1).tolist() == [3.0, 1.0, 19.0]
    assert np.asarray(s).tolist() == [1.0, 0.0, 9.0]
    assert a.counts[1] == 0 and a.counts[-1] == 3
    assert a.content[a.argmax().content].tolist() == [1.0, 2.0]      # integer-array indexing on the device


# ---- exception contract ------------------------------------------------------------------
def test_exceptions(ak):
    J = ak.JaggedArray
    a = std(ak)
    with pytest.raises(TypeError, match="counts must have integer dtype"):
        J.fromcounts(np.array([1.0, 2.0]), [1, 2, 3])
    with pytest.raises(ValueError, match="counts must be a non-negative array"):
        J.fromcounts([1, -1], [1, 2, 3])
    with pytest.raises(ValueError, match="offsets must be monatonically increasing"):
        J.fromoffsets([0, 3, 2], CONTENT).sum()
    with pytest.raises(ValueError, match=r"maximum offset 12 is beyond the length of the content \(10\)"):
        J.fromoffsets([0, 3, 12], CONTENT).sum()
    with pytest.raises(ValueError, match="cannot fit contents of JaggedArray into the given starts and stops arrays"):
        a[J.fromcounts([2, 1, 2, 5], np.ones(10, dtype=bool))]
    with pytest.raises(ValueError, match="jagged array used as index has a different shape"):
        a[J.fromcounts([1], [0])]
    with pytest.raises(IndexError, match="jagged array used as index contains out-of-bounds values"):
        a[J.fromcounts([1, 0, 0, 0], [3])]
    with pytest.raises(TypeError, match=r"jagged index must be boolean \(mask\) or integer \(fancy indexing\)"):
        a[J.fromcounts([1, 0, 0, 0], [0.5])]
    with pytest.raises(ValueError, match=r"cannot broadcast JaggedArray of shape \(4,\) with array of shape \(2,\)"):
        a + np.array([1.0, 2.0])
    with pytest.raises(NotImplementedError, match="in-place operations not supported"):
        np.add(a, a, out=a)
    with pytest.raises(TypeError):
        np.add.reduce(a)
    with pytest.raises(TypeError, match="both arrays must be JaggedArrays"):
        a.cross(np.arange(4))
    with pytest.raises(ValueError, match="both JaggedArrays must have the same length"):
        a.cross(J.fromcounts([1], [1.0]))
    with pytest.raises(ValueError, match="Choosing 0 or 1 items is trivial"):
        a.argchoose(1)
    with pytest.raises(NotImplementedError):
        a.argchoose(6)
    with pytest.raises(ValueError, match="offsets must be a non-negative array"):
        J.fromoffsets([-1, 3], CONTENT)
    with pytest.raises(TypeError, match="invalid index for assigning"):
        a[0] = 1.0
    assert J.fromoffsets([0, 3, 2], CONTENT).valid() is False
    ok = a[J.fromcounts([2, 0, 1, 1], [-1, 0, 1, -5])]       # negative wrap (jagged.py:552-553)
    assert ok.tolist() == [[2.2, 0.0], [], [4.4], [5.5]]


def test_switches(ak):
    J = ak.JaggedArray

    class Unchecked(J):
        check_whole_valid = False

    b = Unchecked.fromoffsets([0, 2, 5], np.arange(5, dtype=np.float32))
    assert b.sum().tolist() == [1.0, 9.0]

    class NoHost(J):
        allow_tonumpy = False

    c = NoHost.fromcounts([1, 1], [1.0, 2.0])
    with pytest.raises(RuntimeError, match="allow_tonumpy is False"):
        c.tolist()
    assert c.sum().tolist() == [1.0, 2.0]          # flat results are plain DeviceArrays


def test_chained_combinatorics_flatten_tuples(ak):      # jagged.py:1241,1271,1294,1350 + table.py:791-820
    J = ak.JaggedArray
    a, b, c = J.fromiter([[1, 2], [], [3]]), J.fromiter([[10], [20], [30, 40]]), J.fromiter([[100, 200], [300], [400]])
    x = a.cross(b).cross(c)                              # literals: the reference run in the build container
    assert x.columns == ["0", "1", "2"]
    assert x.tolist() == [[(1, 10, 100), (1, 10, 200), (2, 10, 100), (2, 10, 200)], [], [(3, 30, 400), (3, 40, 400)]]
    p = a.cross(b).pairs()
    assert p.columns == ["0", "1", "2", "3"] and p.tolist()[0] == [(1, 10, 1, 10), (1, 10, 2, 10), (2, 10, 2, 10)]
    ap = a.pairs().cross(c)
    assert ap.tolist()[0] == [(1, 1, 100), (1, 1, 200), (1, 2, 100), (1, 2, 200), (2, 2, 100), (2, 2, 200)]
    z = J.zip(u=a, v=a * 2)                              # record Tables are not spliced
    zc = z.cross(b)
    assert zc.columns == ["0", "1"] and zc.tolist()[0] == [({"u": 1, "v": 2}, 10), ({"u": 2, "v": 4}, 10)]
    assert (x.i0 + x.i1 + x.i2).tolist() == [[111, 211, 112, 212], [], [433, 443]]
