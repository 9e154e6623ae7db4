#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python oracle/make_golden.py            # rewrites tests/golden/

Each fixture holds seeded inputs (``in_*``) and what awkward0 0.15.5 itself returned for every
hot-path operation (``out_*``).  The fixtures travel to the GPU box (the reference does not), so
both the numpy oracle (``-m "not gpu"``) and the CUDA path (``-m gpu``) are checked against the
real reference's answers.  Test infrastructure only.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import _refloader  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def logical(j):
    """(counts, dense content) of a reference JaggedArray whatever its physical layout."""
    counts = np.asarray(j.counts, dtype=np.int64)
    flat = j.flatten()
    return counts, np.asarray(flat)


def table_cols(t):
    return [np.asarray(t[n]) for n in t.columns]


def run_case(awkward0, name, offsets, content, flat, idx_counts, idx_content, b_counts, b_content,
             threshold, choose_max=5, combos=True):
    J = awkward0.JaggedArray
    out = {}
    out["in_offsets"] = offsets
    out["in_content"] = content
    out["in_flat"] = flat
    out["in_idx_counts"] = idx_counts
    out["in_idx_content"] = idx_content
    out["in_b_counts"] = b_counts
    out["in_b_content"] = b_content
    out["in_threshold"] = np.asarray(threshold)

    a = J.fromoffsets(offsets, content)
    out["out_counts"] = np.asarray(a.counts)
    out["out_parents"] = np.asarray(a.parents)
    li = a.localindex
    out["out_localindex_offsets"] = np.asarray(li.offsets)
    out["out_localindex_content"] = np.asarray(li.content)
    out["out_flatten"] = np.asarray(a.flatten())
    out["out_c2o"] = np.asarray(J.counts2offsets(np.asarray(a.counts)))

    for op in ("sum", "prod", "min", "max", "any", "all"):
        with np.errstate(all="ignore"):
            out["out_" + op] = np.asarray(getattr(a, op)())

    if content.dtype != np.bool_:
        for op in ("argmin", "argmax"):
            c, f = logical(getattr(a, op)())
            out["out_%s_counts" % op] = c
            out["out_%s_flat" % op] = f.astype(np.int64)

        mask = a > threshold
        out["out_mask_offsets"] = np.asarray(mask.offsets)
        out["out_mask_content"] = np.asarray(mask.content)
        sel = a[mask]
        out["out_select_offsets"] = np.asarray(sel.offsets)
        out["out_select_content"] = np.asarray(sel.flatten())

        plus = a + flat
        out["out_plus_offsets"] = np.asarray(plus.offsets)
        out["out_plus_content"] = np.asarray(plus.content)
        times = a * content.dtype.type(threshold)
        out["out_times_content"] = np.asarray(times.content)

    tj = a.tojagged(flat)
    c, f = logical(tj)
    out["out_tojagged_flat"] = f

    idx = J.fromcounts(idx_counts, idx_content)
    got = a[idx]
    c, f = logical(got)
    out["out_gather_counts"] = c
    out["out_gather_flat"] = f

    # argsort and concatenate
    with np.errstate(all="ignore"):
        for label, asc in (("desc", False), ("asc", True)):
            c, f = logical(a.argsort(ascending=asc))
            assert np.array_equal(c, out["out_counts"])
            out["out_argsort_%s_flat" % label] = f.astype(np.int64)
    b = J.fromcounts(b_counts, b_content)
    c, f = logical(J.concatenate([a, b, a]))
    out["out_cat0_counts"], out["out_cat0_flat"] = c, f
    c, f = logical(a.concatenate([b, a], axis=1))
    out["out_cat1_counts"], out["out_cat1_flat"] = c, f

    # pad / fillna; the skewed fixture has a 9000-element event: keep it small there
    fill = {"f": -7.5, "i": 3, "u": 3, "b": True}[content.dtype.kind]
    out["in_fill"] = np.asarray(fill)
    for label, clip in (("pad3", False), ("pad3clip", True)):
        p = a.pad(3, clip=clip)
        out["out_%s_counts" % label] = np.asarray(p.counts)
        out["out_%s_invalid" % label] = np.asarray(p.content.boolmask(maskedwhen=True)) if len(content) else np.ones(3 * len(a), dtype=bool)
        out["out_%s_fill" % label] = np.asarray(p.fillna(fill).flatten())

    # the other constructors and jagged __setitem__ (jagged.py:73-110, 168-242, 789-838)
    par = np.asarray(a.parents)
    if len(par) > 0:            # the reference indexes parents[0] unconditionally (jagged.py:87): no empty case
        rs, re = J.parents2startsstops(par)
        out["out_p2ss_starts"], out["out_p2ss_stops"] = np.asarray(rs), np.asarray(re)
        rs, re = J.parents2startsstops(par[::-1].copy())           # events in reverse order
        out["out_p2ss_rev_starts"], out["out_p2ss_rev_stops"] = np.asarray(rs), np.asarray(re)
        uo, up = J.uniques2offsetsparents(par)
        out["out_u2op_offsets"], out["out_u2op_parents"] = np.asarray(uo), np.asarray(up)
    dense = np.asarray(a.flatten())
    out["out_fromlocalindex_offsets"] = np.asarray(J.fromlocalindex(np.asarray(li.flatten()), dense).offsets)
    out["out_fromfolding7_offsets"] = np.asarray(J.fromfolding(dense, 7).offsets)
    if content.dtype != np.bool_:
        b2 = J.fromoffsets(offsets, content.copy())
        b2[b2 > threshold] = fill
        out["out_setmask_scalar"] = np.asarray(b2.content)
        b2 = J.fromoffsets(offsets, content.copy())
        msk = b2 > threshold
        vals = (np.arange(int(np.asarray(msk.content).sum())) % 5).astype(content.dtype)
        b2[msk] = vals
        out["out_setmask_values"] = np.asarray(b2.content)
        b2 = J.fromoffsets(offsets, content.copy())
        b2[idx] = fill
        out["out_setindex_scalar"] = np.asarray(b2.content)

    if not combos:
        return save(name, out)

    ap = a.argpairs()
    out["out_argpairs_offsets"] = np.asarray(ap.offsets)
    out["out_argpairs_0"], out["out_argpairs_1"] = table_cols(ap.content)
    p = a.pairs()
    out["out_pairs_0"], out["out_pairs_1"] = table_cols(p.content)

    ad = a.argdistincts()
    out["out_argdistincts_offsets"] = np.asarray(ad.offsets)
    out["out_argdistincts_0"], out["out_argdistincts_1"] = table_cols(ad.content)
    d = a.distincts()
    out["out_distincts_0"], out["out_distincts_1"] = table_cols(d.content)

    for k in range(2, choose_max + 1):
        ac = a.argchoose(k)
        out["out_argchoose%d_offsets" % k] = np.asarray(ac.offsets)
        for ci, col in enumerate(table_cols(ac.content)):
            out["out_argchoose%d_%d" % (k, ci)] = col
    ch = a.choose(2)
    out["out_choose2_0"], out["out_choose2_1"] = table_cols(ch.content)

    b = J.fromcounts(b_counts, b_content)
    ax = a.argcross(b)
    out["out_argcross_offsets"] = np.asarray(ax.offsets)
    out["out_argcross_0"], out["out_argcross_1"] = table_cols(ax.content)
    x = a.cross(b)
    out["out_cross_0"], out["out_cross_1"] = table_cols(x.content)

    save(name, out)


def save(name, out):
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def make_idx(rng, counts):
    """A valid jagged integer index: per event 0..3 picks in [-count, count)."""
    n = len(counts)
    want = rng.integers(0, 4, n)
    want[counts == 0] = 0
    total = int(want.sum())
    owner = np.repeat(np.arange(n), want)
    c = counts[owner]
    picks = (rng.integers(0, 1 << 30, total) % (2 * c)) - c
    return want.astype(np.int64), picks.astype(np.int64)


def main():
    awkward0 = _refloader.load()
    os.makedirs(GOLDEN, exist_ok=True)
    rng = np.random.default_rng(20261019)

    def std_case(name, counts, content, lead=0, choose_max=5, threshold=20.0, combos=True):
        counts = np.asarray(counts, dtype=np.int64)
        offsets = np.empty(len(counts) + 1, dtype=np.int64)
        offsets[0] = lead
        offsets[1:] = lead + np.cumsum(counts)
        assert offsets[-1] <= len(content)
        n = len(counts)
        flat = rng.normal(size=n).astype(content.dtype if content.dtype.kind == "f" else np.float64)
        ic, ix = make_idx(rng, counts)
        bc = rng.poisson(2.0, n).astype(np.int64)
        bx = rng.integers(-1000, 1000, int(bc.sum())).astype(np.int32)
        run_case(awkward0, name, offsets, content, flat, ic, ix, bc, bx, threshold, choose_max, combos)

    # 1. SURVEY appendix A vector
    std_case("appendix_a", [3, 0, 2, 4, 1, 0],
             np.array([1.5, np.nan, 1.5, 4.0, 7.0, -2.0, 9.0, 9.0, 0.5, -0.0], dtype=np.float32), threshold=1.0)

    # 2/3. Poisson(5) f32 / f64 (the benchmark distribution)
    counts = rng.poisson(5.0, 1500)
    pt = rng.exponential(25.0, int(counts.sum())).astype(np.float32)
    std_case("poisson5_f32", counts, pt, choose_max=4)
    std_case("poisson5_f64", counts, pt.astype(np.float64), choose_max=3)

    # 4. dimuon-like Poisson(2)
    counts = rng.poisson(2.0, 4000)
    std_case("poisson2_f32", counts, rng.exponential(20.0, int(counts.sum())).astype(np.float32))

    # 5. all events empty
    std_case("all_empty", np.zeros(64, dtype=np.int64), np.empty(0, dtype=np.float32))

    # 6. NaN / inf / -0.0 / ties, with leading unreachable content and trailing empties
    counts = np.concatenate([rng.poisson(3.0, 500), np.zeros(7, dtype=np.int64)])
    vals = rng.integers(-3, 4, int(counts.sum()) + 5).astype(np.float32)
    vals[rng.random(len(vals)) < 0.10] = np.nan
    vals[rng.random(len(vals)) < 0.05] = np.inf
    vals[rng.random(len(vals)) < 0.05] = -np.inf
    vals[rng.random(len(vals)) < 0.10] = -0.0
    std_case("special_f32", counts, vals, lead=5, threshold=0.0)
    std_case("special_f64", counts, vals.astype(np.float64), lead=5, threshold=0.0, choose_max=3)

    # 7. long segments mixed with short ones (exercises the warp / block variants)
    counts = rng.poisson(4.0, 300)
    counts[17] = 9000
    counts[18] = 0
    counts[150] = 700
    counts[299] = 2500
    std_case("skewed_f32", counts, rng.normal(size=int(counts.sum())).astype(np.float32), choose_max=2, threshold=0.5, combos=False)

    # 8. medium segments (mean 40)
    counts = rng.poisson(40.0, 120)
    std_case("poisson40_f64", counts, rng.normal(size=int(counts.sum())), choose_max=2, threshold=0.0)

    # 9. integer / bool content
    counts = rng.poisson(5.0, 1500)
    m = int(counts.sum())
    std_case("ints_i32", counts, rng.integers(-50, 50, m).astype(np.int32), choose_max=2, threshold=3)
    std_case("ints_i64", counts, rng.integers(-2**40, 2**40, m).astype(np.int64), choose_max=2, threshold=0)
    std_case("ints_u8", counts, rng.integers(0, 4, m).astype(np.uint8), choose_max=2, threshold=1)
    std_case("bools", counts, rng.random(m) < 0.7, choose_max=2)

    # 10. single event / single element
    std_case("single", [1], np.array([3.25], dtype=np.float64), threshold=1.0)

    # 11. narrow and unsigned integer content (result dtypes int64 / uint64, wrap-around sums and products)
    counts = rng.poisson(5.0, 400)
    m = int(counts.sum())
    std_case("ints_i8", counts, rng.integers(-128, 128, m).astype(np.int8), choose_max=2, threshold=3)
    std_case("ints_i16", counts, rng.integers(-2**15, 2**15, m).astype(np.int16), choose_max=2, threshold=100)
    std_case("ints_u16", counts, rng.integers(0, 2**16, m).astype(np.uint16), choose_max=2, threshold=30000)
    std_case("ints_u32", counts, rng.integers(0, 2**32, m).astype(np.uint32), choose_max=2, threshold=2**31)
    std_case("ints_u64", counts, rng.integers(0, 2**64, m, dtype=np.uint64), choose_max=2, threshold=2**63)


if __name__ == "__main__":
    main()
