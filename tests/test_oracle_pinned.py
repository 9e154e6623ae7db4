"""Pins oracle/jagged_oracle.py to the reference: (a) golden vectors produced by the unmodified
reference (oracle/make_golden.py), (b) the literal expectations of the reference's own tests."""
import numpy as np
import pytest

from conftest import same
from oracle import jagged_oracle as O

RTOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def test_offsets_machinery(golden):
    g = golden
    off = g["in_offsets"]
    assert same(O.offsets2counts(off), g["out_counts"])
    assert same(O.counts2offsets(g["out_counts"]), g["out_c2o"])
    assert same(O.offsets2parents(off), g["out_parents"])
    lo, lc = O.localindex(off)
    assert same(lo, g["out_localindex_offsets"])
    assert same(lc, g["out_localindex_content"])
    assert same(O.flatten(off, g["in_content"]), g["out_flatten"])
    assert same(O.startsstops2parents(off[:-1], off[1:]), g["out_parents"][:off[-1]]) or len(off) == 1


@pytest.mark.parametrize("op", ["sum", "prod", "min", "max", "any", "all"])
def test_reducers_bit_exact(golden, op):
    # the oracle is sequential like the reference, so even float sums/products are bit-exact here
    g = golden
    with np.errstate(all="ignore"):
        got = O.reduce(g["in_offsets"], g["in_content"], op)
    assert same(got, g["out_" + op]), (g["name"], op)


@pytest.mark.parametrize("op", ["argmin", "argmax"])
def test_argminmax(golden, op):
    g = golden
    if "out_%s_counts" % op not in g:
        pytest.skip("bool content")
    with np.errstate(all="ignore"):
        roff, rflat = O.argminmax(g["in_offsets"], g["in_content"], op == "argmin")
    assert same(O.offsets2counts(roff), g["out_%s_counts" % op])
    assert same(rflat, g["out_%s_flat" % op])


def test_mask_select_and_ufuncs(golden):
    g = golden
    if "out_mask_content" not in g:
        pytest.skip("bool content")
    off, content = g["in_offsets"], g["in_content"]
    thr = g["in_threshold"][()]
    with np.errstate(all="ignore"):
        moff, mask = O.ufunc_broadcast(np.greater, off, content, thr)
    assert same(moff, g["out_mask_offsets"]) and same(mask, g["out_mask_content"])
    noff, sel = O.mask_select(off, content, moff, mask)
    assert same(noff, g["out_select_offsets"]) and same(sel, g["out_select_content"])
    with np.errstate(all="ignore"):
        poff, plus = O.ufunc_broadcast(np.add, off, content, g["in_flat"])
        _, times = O.ufunc_broadcast(np.multiply, off, content, content.dtype.type(thr))
    assert same(poff, g["out_plus_offsets"]) and same(plus, g["out_plus_content"])
    assert same(times, g["out_times_content"])


def test_tojagged_and_gather(golden):
    g = golden
    off, content = g["in_offsets"], g["in_content"]
    assert same(O.tojagged_flat(off, g["in_flat"]), g["out_tojagged_flat"])
    ioff = O.counts2offsets(g["in_idx_counts"])
    goff, got = O.index_gather(off, content, ioff, g["in_idx_content"])
    assert same(O.offsets2counts(goff), g["out_gather_counts"])
    assert same(got, g["out_gather_flat"])


def test_combinatorics(golden):
    g = golden
    if "out_argpairs_offsets" not in g:
        pytest.skip("no combinatorics in this fixture")
    off, content = g["in_offsets"], g["in_content"]
    poff, i, j = O.argpairs(off)
    assert same(poff, g["out_argpairs_offsets"]) and same(i, g["out_argpairs_0"]) and same(j, g["out_argpairs_1"])
    _, v0, v1 = O.pairs(off, content)
    assert same(v0, g["out_pairs_0"]) and same(v1, g["out_pairs_1"])
    poff, i, j = O.argdistincts(off)
    assert same(poff, g["out_argdistincts_offsets"]) and same(i, g["out_argdistincts_0"]) and same(j, g["out_argdistincts_1"])
    _, v0, v1 = O.distincts(off, content)
    assert same(v0, g["out_distincts_0"]) and same(v1, g["out_distincts_1"])
    for k in range(2, 6):
        if "out_argchoose%d_offsets" % k not in g:
            continue
        res = O.argchoose(off, k)
        assert same(res[0], g["out_argchoose%d_offsets" % k])
        for c in range(k):
            assert same(res[1 + c], g["out_argchoose%d_%d" % (k, c)]), (k, c)
    _, v0, v1 = O.choose(off, content, 2)
    assert same(v0, g["out_choose2_0"]) and same(v1, g["out_choose2_1"])
    boff = O.counts2offsets(g["in_b_counts"])
    poff, l, r = O.argcross(off, boff)
    assert same(poff, g["out_argcross_offsets"]) and same(l, g["out_argcross_0"]) and same(r, g["out_argcross_1"])
    _, va, vb = O.cross(off, content, boff, g["in_b_content"])
    assert same(va, g["out_cross_0"]) and same(vb, g["out_cross_1"])


# ---------------------------------------------------------------------------------------------
# literal expectations of the reference's own tests (tests/test_jagged.py, tests/test_issues.py)
# ---------------------------------------------------------------------------------------------

OFF = np.array([0, 3, 3, 5, 10])
CONTENT = np.array([0.0, 1.1, 2.2, 3.3, 4.4, 5.5, 6.6, 7.7, 8.8, 9.9])


@pytest.mark.parametrize("ascending", [False, True])
def test_argsort(golden, ascending):
    g = golden
    with np.errstate(all="ignore"):
        roff, rflat = O.argsort(g["in_offsets"], g["in_content"], ascending)
    assert same(O.offsets2counts(roff), g["out_counts"])
    assert same(rflat, g["out_argsort_%s_flat" % ("asc" if ascending else "desc")])


def test_concatenate(golden):
    g = golden
    a = (g["in_offsets"], g["in_content"])
    b = (O.counts2offsets(g["in_b_counts"]), g["in_b_content"])
    off, content = O.concatenate_axis0([a, b, a])
    assert same(O.offsets2counts(off), g["out_cat0_counts"]) and same(content, g["out_cat0_flat"])
    off, content = O.concatenate_axis1([a, b, a])
    assert same(O.offsets2counts(off), g["out_cat1_counts"]) and same(content, g["out_cat1_flat"])


@pytest.mark.parametrize("clip", [False, True])
def test_pad_fillna(golden, clip):
    g = golden
    label = "pad3clip" if clip else "pad3"
    off = g["in_offsets"]
    fill = g["in_fill"][()]
    poff, data, invalid = O.pad(off[:-1], off[1:], g["in_content"], 3, clip)
    assert same(O.offsets2counts(poff), g["out_%s_counts" % label])
    assert same(invalid, g["out_%s_invalid" % label])
    with np.errstate(all="ignore"):
        foff, filled = O.pad_fillna(off[:-1], off[1:], g["in_content"], 3, fill, clip)
    assert same(filled, g["out_%s_fill" % label])


def test_ref_pad_literals():                    # tests/test_jagged.py:544-559
    def show(starts, stops, content, length, clip):
        off, data, invalid = O.pad(starts, stops, content, length, clip)
        return [[None if invalid[j] else data[j] for j in range(off[i], off[i + 1])] for i in range(len(off) - 1)]
    c = np.array([1.1, 2.2, 3.3, 4.4, 5.5])
    assert show([0, 3, 3], [3, 3, 5], c, 4, True) == [[1.1, 2.2, 3.3, None], [None] * 4, [4.4, 5.5, None, None]]
    assert show([0, 3, 3], [3, 3, 5], c, 4, False) == [[1.1, 2.2, 3.3, None], [None] * 4, [4.4, 5.5, None, None]]
    c = np.array([1.1, 2.2, 3.3, 4.4, 5.5, 6.6, 7.7, 8.8, 9.9])
    s, e = [0, 5, 5, 8], [5, 5, 8, 9]
    assert show(s, e, c, 3, False) == [[1.1, 2.2, 3.3, 4.4, 5.5], [None] * 3, [6.6, 7.7, 8.8], [9.9, None, None]]
    assert show(s, e, c, 3, True) == [[1.1, 2.2, 3.3], [None] * 3, [6.6, 7.7, 8.8], [9.9, None, None]]
    off, filled = O.pad_fillna([0, 3, 3], [3, 3, 5], np.array([1.1, 2.2, 3.3, 4.4, 5.5]), 4, 999)
    assert O.tolist(off, filled) == [[1.1, 2.2, 3.3, 999], [999, 999, 999, 999], [4.4, 5.5, 999, 999]]


def test_constructors_and_setitem(golden):
    g = golden
    off, content = g["in_offsets"], g["in_content"]
    par = g["out_parents"]
    if "out_p2ss_starts" in g:
        s, e = O.parents2startsstops(par)
        assert same(s, g["out_p2ss_starts"]) and same(e, g["out_p2ss_stops"])
        s, e = O.parents2startsstops(par[::-1].copy())
        assert same(s, g["out_p2ss_rev_starts"]) and same(e, g["out_p2ss_rev_stops"])
        uo, up = O.uniques2offsetsparents(par)
        assert same(uo, g["out_u2op_offsets"]) and same(up, g["out_u2op_parents"])
    assert same(O.localindex2offsets(g["out_localindex_content"]), g["out_fromlocalindex_offsets"])
    assert same(O.folding2offsets(int(off[-1] - off[0]), 7), g["out_fromfolding7_offsets"])
    if "out_setmask_scalar" in g:
        fill, thr = g["in_fill"][()], g["in_threshold"][()]
        with np.errstate(all="ignore"):
            moff, mask = O.ufunc_broadcast(np.greater, off, content, thr)
        full_mask = np.zeros(len(content), dtype=bool)
        full_mask[off[0]:off[-1]] = mask
        assert same(O.setitem(off[:-1], off[1:], content, off, full_mask, fill), g["out_setmask_scalar"])
        vals = (np.arange(int(mask.sum())) % 5).astype(content.dtype)
        assert same(O.setitem(off[:-1], off[1:], content, off, full_mask, vals), g["out_setmask_values"])
        ioff = O.counts2offsets(g["in_idx_counts"])
        assert same(O.setitem(off[:-1], off[1:], content, ioff, g["in_idx_content"], fill), g["out_setindex_scalar"])


def test_ref_constructor_literals():            # tests/test_jagged.py:33-35, 50-54
    s, e = O.parents2startsstops([0, 0, 0, 2, 2, 3, 3, 3, 3, 3])
    assert s.tolist() == [0, 0, 3, 5] and e.tolist() == [3, 0, 5, 10]
    o, p = O.uniques2offsetsparents([9, 9, 9, 8, 8, 7, 7, 7, 7, 7])
    assert o.tolist() == [0, 3, 5, 10] and p.tolist() == [0, 0, 0, 1, 1, 2, 2, 2, 2, 2]
    assert O.localindex2offsets([0, 1, 0, 0, 0, 1, 2, 0, 1, 0]).tolist() == [0, 2, 3, 4, 7, 9, 10]
    with pytest.raises(ValueError):
        O.localindex2offsets([0, 2])


def test_ref_setitem_literals():                # tests/test_jagged.py:609-657
    for starts, stops, content in (([0, 1, 1], [1, 1, 3], [1.1, 2.2, 3.3]), ([3, 4, 1], [4, 4, 3], [0.0, 2.2, 3.3, 1.1, 4.4])):
        def show(c):
            return [c[a:b].tolist() for a, b in zip(starts, stops)]
        woff = np.array([0, 1, 1, 3])
        c = O.setitem(starts, stops, np.array(content), woff, np.array([True, True, True]), np.array([4.4, 5.5, 6.6]))
        assert show(c) == [[4.4], [], [5.5, 6.6]]
        c = O.setitem(starts, stops, c, woff, np.array([False, True, True]), np.array([7.7, 8.8]))
        assert show(c) == [[4.4], [], [7.7, 8.8]]
        c = O.setitem(starts, stops, c, woff, np.array([False, True, False]), 9.9)
        assert show(c) == [[4.4], [], [9.9, 8.8]]
        c = O.setitem(starts, stops, np.array(content), woff, np.array([0, 0, 1]), np.array([4.4, 5.5, 6.6]))
        assert show(c) == [[4.4], [], [5.5, 6.6]]
        c = O.setitem(starts, stops, c, np.array([0, 0, 0, 2]), np.array([0, 1]), np.array([7.7, 8.8]))
        assert show(c) == [[4.4], [], [7.7, 8.8]]
        c = O.setitem(starts, stops, c, np.array([0, 0, 0, 1]), np.array([0]), 9.9)
        assert show(c) == [[4.4], [], [9.9, 8.8]]


def test_ref_concatenate_literals():           # tests/test_jagged.py:451-487
    lst = [[1, 2, 3], [], [4, 5], [6, 7], [8], [], [9], [10, 11], [12]]

    def build(rows):
        counts = np.array([len(r) for r in rows], dtype=np.int64)
        return O.counts2offsets(counts), np.array([x for r in rows for x in r], dtype=np.int64)
    off, content = O.concatenate_axis0([build(lst[:3]), build(lst[3:6]), build(lst[6:])])
    assert O.tolist(off, content) == lst
    c = np.array([0.0, 1.1, 2.2, 3.3, 4.4, 5.5, 6.6, 7.7, 8.8, 9.9])
    d = np.array([6.5, 7.6, 8.7, 9.8, 10.9, 4.3, 5.4, 1., 2.1, 3.2])
    off, content = O.concatenate_axis1([(np.array([0, 0, 3, 3, 5, 10]), c), (np.array([0, 0, 5, 7, 7, 10]), d)])
    assert O.tolist(off, content) == [[], [0., 1.1, 2.2, 6.5, 7.6, 8.7, 9.8, 10.9], [4.3, 5.4], [3.3, 4.4],
                                      [5.5, 6.6, 7.7, 8.8, 9.9, 1., 2.1, 3.2]]
    f, t = np.zeros(3, dtype=bool), np.ones(3, dtype=bool)
    assert O.concatenate_axis1([(np.array([0, 3]), f), (np.array([0, 3]), t)])[1].dtype == np.dtype(bool)
    for dt in (np.int32, np.int64, np.float32, np.float64):
        one = (np.array([0, 1]), np.ones(1, dtype=dt))
        assert O.concatenate_axis1([one, one])[1].dtype == np.dtype(dt)


def test_ref_argsort_literals():                # tests/test_jagged.py:604-607 (without the masked event)
    off = np.array([0, 3, 6, 10, 11])
    c = np.array([2., 3., 1., 4., -np.inf, 5., np.inf, 4., np.nan, -np.inf, np.nan])
    assert O.tolist(*O.argsort(off, c)) == [[1, 0, 2], [2, 0, 1], [0, 1, 3, 2], [0]]
    assert O.tolist(*O.argsort(off, c, True)) == [[2, 0, 1], [1, 0, 2], [3, 1, 0, 2], [0]]
    off = np.array([0, 3, 3, 5, 10])
    c = np.array([3.3, 1.1, 2.2, 5.5, 4.4, 9.9, 6.6, 7.7, 8.8, 0.0])
    assert O.tolist(*O.argsort(off, c)) == [[0, 2, 1], [], [0, 1], [0, 3, 2, 1, 4]]
    assert O.tolist(*O.argsort(off, c, True)) == [[1, 2, 0], [], [1, 0], [4, 1, 2, 3, 0]]
    c[[1, 6]] = np.nan
    assert O.tolist(*O.argsort(off, c)) == [[0, 2, 1], [], [0, 1], [0, 3, 2, 4, 1]]
    assert O.tolist(*O.argsort(off, c, True)) == [[2, 0, 1], [], [1, 0], [4, 2, 3, 0, 1]]


def test_ref_sum_prod_min_max_literals():
    # tests/test_jagged.py:384-386, 398-400, 425-427, 438-440
    assert O.reduce(OFF, CONTENT, "sum").tolist() == [3.3000000000000003, 0.0, 7.7, 38.5]
    assert O.reduce(OFF, CONTENT, "prod").tolist() == [0.0, 1.0, 14.52, 24350.911200000002]
    assert O.reduce(OFF, CONTENT, "min").tolist() == [0.0, np.inf, 3.3, 5.5]
    assert O.reduce(OFF, CONTENT, "max").tolist() == [2.2, -np.inf, 4.4, 9.9]


def test_ref_argminmax_literals():
    # tests/test_jagged.py:411-423
    o, f = O.argminmax(OFF, CONTENT, True)
    assert O.tolist(o, f) == [[0], [], [0], [0]]
    o, f = O.argminmax(OFF, CONTENT, False)
    assert O.tolist(o, f) == [[2], [], [1], [4]]


def test_ref_ufunc_literals():
    # tests/test_jagged.py:151-154
    o, c = O.ufunc_broadcast(np.add, OFF, CONTENT, 100)
    assert O.tolist(o, c) == [[100.0, 101.1, 102.2], [], [103.3, 104.4], [105.5, 106.6, 107.7, 108.8, 109.9]]
    o, c = O.ufunc_broadcast(np.add, OFF, CONTENT, np.array([100, 200, 300, 400]))
    assert O.tolist(o, c) == [[100.0, 101.1, 102.2], [], [303.3, 304.4], [405.5, 406.6, 407.7, 408.8, 409.9]]


def test_ref_tojagged_literal():
    # tests/test_jagged.py:56-60
    off = np.array([0, 1, 3, 3])
    assert O.tolist(off, O.tojagged_flat(off, np.array([3, 4, 5]) + 1)) == [[4], [5, 5], []]


def test_ref_localindex_parents_literals():
    # tests/test_jagged.py:592-602
    off = np.array([0, 5, 5, 8, 9])
    lo, lc = O.localindex(off)
    assert O.tolist(lo, lc) == [[0, 1, 2, 3, 4], [], [0, 1, 2], [0]]
    assert O.offsets2parents(off).tolist() == [0, 0, 0, 0, 0, 2, 2, 2, 3]
    # b = a[[False, True, False, True]] -> starts [5, 8], stops [5, 9]
    assert O.startsstops2parents(np.array([5, 8]), np.array([5, 9])).tolist() == [-1, -1, -1, -1, -1, -1, -1, -1, 1]


def test_ref_pairs_cross_distincts_loops():
    # tests/test_jagged.py:214-252 (sizes 0..49 against python loops)
    for i in range(50):
        off = np.array([0, 0, 1, 1 + i, 1 + i])
        poff, l, r = O.argpairs(off)
        assert O.offsets2counts(poff).tolist() == [0, 1, i * (i + 1) // 2, 0]
        left = sum([[x] * (i - x) for x in range(i)], [])
        right = sum([list(range(x, i)) for x in range(i)], [])
        assert l[poff[2]:poff[3]].tolist() == left and r[poff[2]:poff[3]].tolist() == right
        doff, l, r = O.argdistincts(off)
        assert l[doff[2]:doff[3]].tolist() == [x for x, y in zip(left, right) if x != y]
        assert r[doff[2]:doff[3]].tolist() == [y for x, y in zip(left, right) if x != y]
    for i in range(10):
        for j in range(5):
            xoff, l, r = O.argcross(np.array([0, 0, 1, 1 + i, 1 + i]), np.array([0, 0, 1, 1 + j, 2 + j]))
            assert O.offsets2counts(xoff).tolist() == [0, 1, i * j, 0]
            assert l[xoff[2]:xoff[3]].tolist() == np.repeat(range(i), j).tolist()
            assert r[xoff[2]:xoff[3]].tolist() == np.tile(range(j), i).tolist()


def test_ref_argchoose_loops():
    # tests/test_jagged.py:254-338 : colex order against itertools
    import itertools
    for k in (2, 3, 4, 5):
        for n in range(0, 12):
            off = np.array([0, 0, 1, 1 + n, 1 + n])
            res = O.argchoose(off, k)
            poff, cols = res[0], res[1:]
            lo, hi = poff[2], poff[3]
            want = sorted(itertools.combinations(range(n), k), key=lambda t: tuple(reversed(t)))
            got = list(zip(*[c[lo:hi].tolist() for c in cols]))
            assert got == want, (k, n)


def test_ref_issue_mask_on_offset_content():
    # tests/test_issues.py:16-25 style: content starts at offset 2; empty events in mask+sum
    off = np.array([2, 4, 4, 7])
    content = np.arange(10, dtype=np.float64)
    moff, mask = O.ufunc_broadcast(np.greater, off, content, 3)
    noff, sel = O.mask_select(off, content, moff, mask)
    assert O.tolist(noff, sel) == [[], [], [4.0, 5.0, 6.0]]
    assert O.reduce(noff, sel, "sum").tolist() == [0.0, 0.0, 15.0]


def test_ref_exceptions():
    with pytest.raises(TypeError, match="counts must have integer dtype"):
        O.check_counts(np.array([1.0]))
    with pytest.raises(ValueError, match="counts must be a non-negative array"):
        O.check_counts(np.array([1, -1]))
    with pytest.raises(ValueError, match="monatonically"):
        O.check_offsets(np.array([0, 3, 2]), 10)
    with pytest.raises(ValueError, match="beyond the length"):
        O.check_offsets(np.array([0, 3, 12]), 10)
    with pytest.raises(IndexError, match="out-of-bounds"):
        O.index_gather(OFF, CONTENT, np.array([0, 1, 1, 1, 1]), np.array([3]))
    with pytest.raises(ValueError, match="cannot fit contents"):
        O.mask_select(OFF, CONTENT, np.array([0, 2, 2, 2, 2]), np.array([True, False]))
    with pytest.raises(ValueError, match="cannot broadcast JaggedArray"):
        O.ufunc_broadcast(np.add, OFF, CONTENT, np.array([1.0, 2.0]))
