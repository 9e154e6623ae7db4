"""CPU oracle for the awkward0 JaggedArray hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a plain-numpy *restatement* of the algorithms the reference
(scikit-hep/awkward-0.x, ``awkward0/array/jagged.py`` and the reducer front-ends of
``awkward0/array/base.py``) uses on the JaggedArray hot path.  Every function cites the
reference lines it follows.  It exists so that the CUDA product path can be checked
bit-for-bit (or to the stated rtol) on a box where ``/root/reference`` does not exist.

Rules (see DESIGN.md "Oracle"):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
    ``--impl reference`` legs may import this module;
  * the shipped package (``awkward-0.x_b200/``) never imports it and has no CPU fallback;
  * parity is PINNED: ``tests/test_oracle_pinned.py`` checks every function here against
    (a) the literal expected values of the reference's own tests (tests/test_jagged.py,
    tests/test_issues.py, README examples) and (b) golden vectors that
    ``oracle/make_golden.py`` produced by running the unmodified reference in the build
    container (committed under ``tests/golden/``).

The functions are deliberately *functional* (arrays in, arrays out) and use the same numpy
primitives as the reference (cumsum / repeat / ufunc.reduceat / fancy indexing) so that timing
them is an honest stand-in for the reference's CPU cost ("kind": "port" in bench.py).

A jagged array is passed around as ``(offsets, content)`` when it is dense ("compact") or as
``(starts, stops, content)`` in the general case; ``INDEX`` is the reference's INDEXTYPE
(int64, base.py:44).
"""

import numpy as np

INDEX = np.dtype(np.int64)      # base.py:44  INDEXTYPE
BOOL = np.dtype(np.bool_)       # base.py:48  BOOLTYPE


# --------------------------------------------------------------------------------------
# offsets / parents machinery
# --------------------------------------------------------------------------------------

def counts2offsets(counts):
    """jagged.py:42-47 -- offsets[0] = 0, offsets[1:] = cumsum(counts); always int64."""
    counts = np.asarray(counts)
    out = np.empty(len(counts) + 1, dtype=INDEX)
    out[0] = 0
    np.cumsum(counts, out=out[1:])
    return out


def check_counts(counts):
    """jagged.py:156-161 -- the TypeError / ValueError fromcounts raises."""
    counts = np.asarray(counts)
    if not issubclass(counts.dtype.type, np.integer):
        raise TypeError("counts must have integer dtype")
    if (counts < 0).any():
        raise ValueError("counts must be a non-negative array")
    return counts


def check_offsets(offsets, content_len):
    """jagged.py:120-127 and 469-477 -- property checks, then the whole-array checks."""
    offsets = np.asarray(offsets)
    if not issubclass(offsets.dtype.type, np.integer):
        raise TypeError("offsets must have integer dtype")
    if offsets.ndim != 1:
        raise ValueError("offsets must be a one-dimensional array")
    if len(offsets) == 0:
        raise ValueError("offsets must be a non-empty array")
    if (offsets < 0).any():
        raise ValueError("offsets must be a non-negative array")
    if not (offsets[1:] >= offsets[:-1]).all():
        raise ValueError("offsets must be monatonically increasing")
    if offsets.max() > content_len:
        raise ValueError("maximum offset {0} is beyond the length of the content ({1})".format(offsets.max(), content_len))
    return offsets


def offsets2counts(offsets):
    """jagged.py:383-388 -- counts = stops - starts (dtype of the offsets is kept)."""
    offsets = np.asarray(offsets)
    return offsets[1:] - offsets[:-1]


def offsets2parents(offsets):
    """jagged.py:49-54 -- repeat(arange(-1, N), diff([0, offsets])): content below offsets[0] gets -1."""
    offsets = np.asarray(offsets)
    reps = np.diff(np.concatenate([np.zeros(1, dtype=INDEX), offsets.astype(INDEX)]))
    return np.repeat(np.arange(-1, len(offsets) - 1, dtype=INDEX), reps)


def startsstops2parents(starts, stops):
    """jagged.py:56-71 -- general (non-dense) layout: unreachable content gets -1."""
    starts = np.asarray(starts).reshape(-1)
    stops = np.asarray(stops).reshape(-1)
    out = np.full(stops.max() if len(stops) else 0, -1, dtype=INDEX)
    counts = stops - starts
    owner = np.repeat(np.arange(len(starts), dtype=INDEX), counts)
    dense = counts2offsets(counts)
    where = np.arange(dense[-1], dtype=INDEX) - np.repeat(dense[:-1], counts) + starts[owner]
    out[where] = owner
    return out


def localindex(offsets):
    """jagged.py:430-437 -- (rebased offsets, arange(o[0], o[-1]) - offsets[parents>=0])."""
    offsets = np.asarray(offsets)
    if len(offsets) <= 1:
        return np.zeros(1, dtype=INDEX), np.empty(0, dtype=INDEX)
    par = offsets2parents(offsets)
    out = np.arange(offsets[0], offsets[-1], dtype=INDEX)
    out -= offsets[par[par >= 0]]
    return (offsets - offsets[0]).astype(offsets.dtype), out


def is_offsets_compatible(starts, stops):
    """jagged.py:358,1369-1371 -- starts[1:] == stops[:-1]."""
    starts, stops = np.asarray(starts), np.asarray(stops)
    return starts.ndim == 1 and np.array_equal(starts[1:], stops[:len(starts)][:-1])


def compact(starts, stops, content):
    """jagged.py:1386-1398 + 883-942 -- re-lay content densely; returns (offsets, content)."""
    starts = np.asarray(starts)
    stops = np.asarray(stops)[:len(starts)]
    content = np.asarray(content)
    counts = stops - starts
    offsets = counts2offsets(counts)
    owner = np.repeat(np.arange(len(starts), dtype=INDEX), counts)
    local = np.arange(offsets[-1], dtype=INDEX) - offsets[:-1][owner]
    return offsets, content[starts[owner] + local]


def flatten(offsets, content):
    """jagged.py:1418-1422 -- dense layout: the view content[starts[0]:stops[-1]]."""
    offsets = np.asarray(offsets)
    content = np.asarray(content)
    if len(offsets) <= 1:
        return content[0:0]
    return content[offsets[0]:offsets[-1]]


# --------------------------------------------------------------------------------------
# segmented reductions
# --------------------------------------------------------------------------------------

_REDUCERS = {
    # name: (ufunc, identity, forced dtype)       base.py:193-215
    "sum":  (np.add, 0, None),
    "prod": (np.multiply, 1, None),
    "min":  (np.minimum, np.inf, None),
    "max":  (np.maximum, -np.inf, None),
    "any":  (np.bitwise_or, False, BOOL),
    "all":  (np.bitwise_and, True, BOOL),
}


def reduce(offsets, content, op):
    """jagged.py:1459-1565 (1-D content, dense offsets) via base.py:193-215.

    NaNs become the identity before reducing (:1524-1528); bool content is promoted to the
    identity's Python type (:1530-1531); +-inf identities become iinfo.max/min for integer
    output (:1537-1547); trailing empties are padded (:1552-1561) and every empty event is
    overwritten with the identity (:1563).
    """
    ufunc, identity, dtype = _REDUCERS[op]
    offsets = np.asarray(offsets)
    content = np.asarray(content)
    nevents = len(offsets) - 1

    if issubclass(content.dtype.type, np.floating):
        nan = np.isnan(content)
        if nan.any():
            content = content.copy()
            content[nan] = identity

    if dtype is None and issubclass(content.dtype.type, np.bool_):
        dtype = np.dtype(type(identity))
    if dtype is None:
        dtype = content.dtype
    else:
        dtype = np.dtype(dtype)
        content = content.astype(dtype)

    if identity == np.inf:
        if issubclass(dtype.type, np.bool_):
            identity = True
        elif issubclass(dtype.type, np.integer):
            identity = np.iinfo(dtype.type).max
    elif identity == -np.inf:
        if issubclass(dtype.type, np.bool_):
            identity = False
        elif issubclass(dtype.type, np.integer):
            identity = np.iinfo(dtype.type).min

    out = np.empty(nevents, dtype=dtype)
    if nevents != 0:
        live = offsets[offsets < len(content)]
        if len(live) != 0:
            got = ufunc.reduceat(content, live)[:nevents]
        else:
            got = np.empty(0, dtype=dtype)
        if len(got) < nevents:
            padded = np.empty(nevents, dtype=got.dtype)
            padded[:len(got)] = got
            got = padded
        out = got
        out[offsets[:-1] == offsets[1:]] = identity
    return out


def argminmax(offsets, content, ismin):
    """jagged.py:1581-1601 -- LOGICAL value of argmin/argmax, returned dense.

    The reference computes ``localindex[self.max() == self]`` and then truncates every
    non-empty event of the result to its first element (:1597-1600), leaving a non-dense
    starts/stops over a content that still holds all ties.  Parity is defined on the logical
    value, so this returns ``(offsets_of_result, flat_indices)`` with 0 or 1 index per event.
    """
    offsets = np.asarray(offsets)
    content = np.asarray(content)
    nevents = len(offsets) - 1
    if nevents == 0 or len(content) == 0:
        return np.zeros(nevents + 1, dtype=INDEX), np.empty(0, dtype=INDEX)
    best = reduce(offsets, content, "min" if ismin else "max")
    par = offsets2parents(offsets)
    inside = par >= 0
    dense_content = content[offsets[0]:offsets[-1]]
    par_dense = par[inside]
    hit = best[par_dense] == dense_content                       # :1592/1594 via __array_ufunc__
    _, local = localindex(offsets)
    hitcounts = reduce_counts_of_mask(offsets - offsets[0], hit)
    hitoffsets = counts2offsets(hitcounts)
    picked = local[hit]
    has = hitcounts > 0
    first = picked[hitoffsets[:-1][has]]                          # first tie only (:1597-1600)
    return counts2offsets(has.astype(INDEX)), first


def reduce_counts_of_mask(offsets, mask):
    """jagged.py:574-575 -- per-event popcount of a jagged boolean (astype(int64) then sum)."""
    return reduce(offsets, np.asarray(mask).astype(INDEX), "sum")


# --------------------------------------------------------------------------------------
# jagged __getitem__
# --------------------------------------------------------------------------------------

def mask_select(offsets, content, mask_offsets, mask):
    """jagged.py:562-589 -- a[jagged_bool]; returns (new offsets, selected content).

    The mask must have the same per-event counts (the _tojagged re-alignment at :573 raises the
    ValueError of :911-912 otherwise).
    """
    offsets = np.asarray(offsets)
    mask_offsets = np.asarray(mask_offsets)
    content = np.asarray(content)
    mask = np.asarray(mask)
    if not (issubclass(mask.dtype.type, np.bool_) or issubclass(mask.dtype.type, np.integer)):
        raise TypeError("jagged index must be boolean (mask) or integer (fancy indexing)")
    if len(mask_offsets) != len(offsets) or not np.array_equal(offsets2counts(mask_offsets), offsets2counts(offsets)):
        raise ValueError("cannot fit contents of JaggedArray into the given starts and stops arrays")
    dense_mask = mask[mask_offsets[0]:mask_offsets[-1]] if len(mask_offsets) > 1 else mask[0:0]
    dense_content = flatten(offsets, content)
    rebased = offsets - (offsets[0] if len(offsets) else 0)
    newoffsets = counts2offsets(reduce_counts_of_mask(rebased, dense_mask))
    return newoffsets, dense_content[dense_mask.astype(BOOL)]


def index_gather(offsets, content, index_offsets, index):
    """jagged.py:541-560 -- a[jagged_int]; returns (dense offsets of the index, gathered content).

    Negative indices wrap by the event's count (:552-553); anything still outside [0, count)
    raises IndexError (:555-556); a different number of events raises ValueError (:542-543).
    """
    offsets = np.asarray(offsets)
    index_offsets = np.asarray(index_offsets)
    content = np.asarray(content)
    index = np.asarray(index)
    if len(index_offsets) != len(offsets):
        raise ValueError("jagged array used as index has a different shape {0} from the jagged array it is selecting from {1}".format((len(index_offsets) - 1,), (len(offsets) - 1,)))
    icounts = offsets2counts(index_offsets)
    headoffsets = counts2offsets(icounts)
    dense_index = np.array(index[index_offsets[0]:index_offsets[-1]] if len(index_offsets) > 1 else index[0:0], dtype=INDEX)
    owner = np.repeat(np.arange(len(icounts), dtype=INDEX), icounts)
    counts = offsets2counts(offsets)[owner]
    neg = dense_index < 0
    dense_index[neg] += counts[neg]
    if not np.bitwise_and(0 <= dense_index, dense_index < counts).all():
        raise IndexError("jagged array used as index contains out-of-bounds values")
    dense_index += offsets[:-1][owner]
    return headoffsets, content[dense_index]


# --------------------------------------------------------------------------------------
# broadcasting
# --------------------------------------------------------------------------------------

def tojagged_flat(offsets, flat):
    """jagged.py:866-875 -- flat[N] -> flat[parents]; logical (dense) content is returned."""
    offsets = np.asarray(offsets)
    flat = np.asarray(flat)
    if flat.ndim == 0 or (flat.ndim == 1 and flat.shape[0] == 1):
        return np.full(int(offsets[-1] - offsets[0]) if len(offsets) > 1 else 0, flat.reshape(-1)[0], dtype=flat.dtype)
    if len(flat) != len(offsets) - 1:
        raise ValueError("cannot broadcast AwkwardArray to match JaggedArray with a different length")
    par = offsets2parents(offsets)
    return flat[par[par >= 0]]


def ufunc_broadcast(ufunc, offsets, content, other):
    """jagged.py:944-1051 -- ufunc(jagged, scalar | flat[N] | dense content of same counts).

    Returns (offsets rebased to 0 by the trailing fromcounts at :1045-1051, result content).
    """
    offsets = np.asarray(offsets)
    dense = flatten(offsets, np.asarray(content))
    other = np.asarray(other)
    if other.ndim == 1:
        if other.shape != (len(offsets) - 1,):
            raise ValueError("cannot broadcast JaggedArray of shape {0} with array of shape {1}".format((len(offsets) - 1,), other.shape))
        par = offsets2parents(offsets)
        other = other[par[par >= 0]]
    result = ufunc(dense, other)
    return counts2offsets(offsets2counts(offsets)), result


# --------------------------------------------------------------------------------------
# combinatorics
# --------------------------------------------------------------------------------------

def argpairs(offsets):
    """jagged.py:1070-1091,1279-1287 -- all (i <= j) per event, lexicographic, local indices."""
    counts = offsets2counts(np.asarray(offsets)).astype(INDEX)
    npairs = counts * (counts + 1) >> 1
    poff = counts2offsets(npairs)
    par = offsets2parents(poff)
    n = counts[par]
    k = np.arange(poff[-1], dtype=INDEX) - poff[par]
    b = 2 * n + 1
    i = np.floor((b - np.sqrt(b * b - 8 * k)) / 2).astype(INDEX)
    j = k - n * i + (i * (i + 1) >> 1)
    return poff, i, j


def argdistincts(offsets):
    """jagged.py:1093-1126 -- all (i < j) per event, lexicographic, local indices."""
    counts = offsets2counts(np.asarray(offsets)).astype(INDEX)
    ncomb = counts * (counts - 1) // 2
    poff = counts2offsets(ncomb)
    par = offsets2parents(poff)
    k = np.arange(poff[-1], dtype=INDEX) - poff[par]
    n = counts[par]
    b = 2 * n - 1
    i = np.floor((b - np.sqrt(b * b - 8 * k)) / 2).astype(INDEX)
    j = k + i * (i - b + 2) // 2 + 1
    return poff, i, j


def _root2(a):
    return np.floor((1 + np.sqrt(8 * a + 1)) / 2).astype(INDEX)                     # jagged.py:1131-1132


def _root3(a):
    out = 2 * np.ones(a.shape, dtype=INDEX)                                        # jagged.py:1134-1140
    pos = a > 0
    rad = np.power(np.sqrt(3) * np.sqrt(243 * a[pos] ** 2 - 1) + 27 * a[pos], 1. / 3)
    out[pos] = np.floor(np.power(3, -2. / 3) * rad + np.power(3, -1. / 3) / rad + 1 + 1e-12)
    return out


def _root4(a):
    return np.floor((np.sqrt(4 * np.sqrt(24 * a + 1) + 5) + 3) / 2).astype(INDEX)    # jagged.py:1142-1144


def argchoose(offsets, k):
    """jagged.py:1128-1230 -- k-combinations (k = 2..5) in colexicographic order, local indices."""
    if k > 5:
        raise NotImplementedError
    if k <= 1:
        raise ValueError("Choosing 0 or 1 items is trivial")
    offsets = np.asarray(offsets)
    c = offsets2counts(offsets).astype(INDEX)
    if k == 5:
        ncomb = c * (c - 1) * (c - 2) * (c - 3) * (c - 4) // 120
    elif k == 4:
        ncomb = c * (c - 1) * (c - 2) * (c - 3) // 24
    elif k == 3:
        ncomb = c * (c - 1) * (c - 2) // 6
    else:
        ncomb = c * (c - 1) // 2
    _, loc = localindex(offsets)
    poff = counts2offsets(ncomb)
    par = offsets2parents(poff)
    kk = np.arange(poff[-1], dtype=INDEX) - poff[par]
    cols = []
    if k == 5:
        i5 = np.repeat(loc, loc * (loc - 1) * (loc - 2) * (loc - 3) // 24)
        kk = kk - i5 * (i5 - 1) * (i5 - 2) * (i5 - 3) * (i5 - 4) // 120
        cols.append(i5)
    if k >= 4:
        i4 = np.repeat(loc, loc * (loc - 1) * (loc - 2) // 6) if k == 4 else _root4(kk)
        kk = kk - i4 * (i4 - 1) * (i4 - 2) * (i4 - 3) // 24
        cols.append(i4)
    if k >= 3:
        i3 = np.repeat(loc, loc * (loc - 1) // 2) if k == 3 else _root3(kk)
        kk = kk - i3 * (i3 - 1) * (i3 - 2) // 6
        cols.append(i3)
    i2 = np.repeat(loc, loc) if k == 2 else _root2(kk)
    kk = kk - i2 * (i2 - 1) // 2
    cols.append(i2)
    cols.append(kk)
    return (poff,) + tuple(reversed(cols))


def argcross(offsets_a, offsets_b):
    """jagged.py:1302-1337 -- per-event Cartesian product, left index slow, local indices."""
    offsets_a = np.asarray(offsets_a)
    offsets_b = np.asarray(offsets_b)
    if len(offsets_a) != len(offsets_b):
        raise ValueError("both JaggedArrays must have the same length")
    ca = offsets2counts(offsets_a).astype(INDEX)
    cb = offsets2counts(offsets_b).astype(INDEX)
    poff = counts2offsets(ca * cb)
    par = offsets2parents(poff)
    nb = cb[par]
    k = np.arange(poff[-1], dtype=INDEX) - poff[par]
    left = k // nb
    right = k - nb * left
    return poff, left, right


def gather_columns(offsets, content, poff, *local_cols):
    """jagged.py:1294,1271,1240,1350 -- content[start_of_event + local index] for each column."""
    offsets = np.asarray(offsets)
    content = np.asarray(content)
    par = offsets2parents(poff)
    base = offsets[:-1][par]
    return tuple(content[base + col] for col in local_cols)


def pairs(offsets, content):
    poff, i, j = argpairs(offsets)
    return (poff,) + gather_columns(offsets, content, poff, i, j)


def distincts(offsets, content):
    poff, i, j = argdistincts(offsets)
    return (poff,) + gather_columns(offsets, content, poff, i, j)


def choose(offsets, content, k):
    res = argchoose(offsets, k)
    return (res[0],) + gather_columns(offsets, content, res[0], *res[1:])


def cross(offsets_a, content_a, offsets_b, content_b):
    poff, left, right = argcross(offsets_a, offsets_b)
    (va,) = gather_columns(offsets_a, content_a, poff, left)
    (vb,) = gather_columns(offsets_b, content_b, poff, right)
    return poff, va, vb


# --------------------------------------------------------------------------------------
# helpers for tests
# --------------------------------------------------------------------------------------

def concatenate_axis0(parts):
    """jagged.py:1633-1649 (via base.py:567-605) -- append the events of dense arrays [(offsets, content), ...].

    The reference concatenates starts / stops / content and shifts the starts and stops of every later array by
    the length of the content in front of it; on dense inputs that is one offsets array.  Content dtypes promote
    like numpy.concatenate.
    """
    offs, contents, base = [np.zeros(1, dtype=INDEX)], [], 0
    for offsets, content in parts:
        offsets = np.asarray(offsets, dtype=INDEX)
        content = np.asarray(content)[offsets[0]:offsets[-1]]
        offs.append(offsets[1:] - offsets[0] + base)
        contents.append(content)
        base += len(content)
    return np.concatenate(offs), np.concatenate(contents)


def concatenate_axis1(parts):
    """jagged.py:1651-1733 -- event i of the result is event i of every input, one after the other.

    The reference stacks the counts (vstack, transpose, flatten), scans them, and fills the content through one
    boolean mask per input (+1 at the slot starts, -1 at the slot stops, cumsum); the slots are contiguous so
    this is a per-event copy.  dtype: numpy promotion of the non-empty inputs, bool only if all are bool
    (:1690-1699); all inputs empty -> dtype of the first.
    """
    n = len(parts[0][0]) - 1
    if any(len(o) - 1 != n for o, _ in parts):
        raise ValueError("cannot concatenate JaggedArrays of different lengths with axis=1")
    dense = [(np.asarray(o, dtype=INDEX), np.asarray(c)) for o, c in parts]
    counts = np.vstack([offsets2counts(o) for o, _ in dense])
    slots = counts2offsets(counts.T.flatten())
    flats = [c[o[0]:o[-1]] for o, c in dense]
    nonempty = [f for f in flats if len(f)]
    if not nonempty:
        dtype = flats[0].dtype
    elif all(f.dtype == np.dtype(bool) for f in flats):
        dtype = np.dtype(bool)
    else:
        dtype = np.array([f[0] for f in nonempty]).dtype
    content = np.zeros(slots[-1], dtype=dtype)
    for k, (o, _) in enumerate(dense):
        starts_k = slots[k:-1:len(dense)]
        c = counts[k]
        dst = np.repeat(starts_k - (o[:-1] - o[0]), c) + np.arange(len(flats[k]), dtype=INDEX)
        content[dst] = flats[k]
    return slots[::len(dense)].copy(), content


def argsort(offsets, content, ascending=False):
    """jagged.py:1603-1631 -- per-event argsort as LOCAL indices, same counts as the input.

    The reference repeatedly takes ``reducer(tmp) == tmp`` (every element equal to the event's current extreme),
    appends their local indices in position order and removes them; NaN compares unequal to everything, so NaN
    stays until only NaN is left (:1618-1620) and is then emitted in position order.  Logically that is a stable
    sort by value (descending unless ``ascending``) with ties -- including -0.0 == 0.0 -- in position order and
    NaN last in both directions.
    """
    offsets = np.asarray(offsets, dtype=INDEX)
    content = np.asarray(content)
    out = np.empty(offsets[-1] - offsets[0], dtype=INDEX)
    base = offsets[0]
    for i in range(len(offsets) - 1):
        x = content[offsets[i]:offsets[i + 1]]
        if len(x) == 0:
            continue
        nan = np.isnan(x) if x.dtype.kind == "f" else np.zeros(len(x), dtype=bool)
        good, bad = np.nonzero(~nan)[0], np.nonzero(nan)[0]
        v = x[good]
        if x.dtype.kind == "b":
            v = v.astype(np.int8)
        if ascending:
            order = np.argsort(v, kind="stable")
        elif x.dtype.kind == "u":
            order = np.argsort(~v, kind="stable")            # descending without leaving the unsigned type
        elif x.dtype.kind == "i":
            order = np.argsort(-1 - v, kind="stable")         # exact for the most negative value too
        else:
            order = np.argsort(-v, kind="stable")
        out[offsets[i] - base:offsets[i + 1] - base] = np.concatenate([good[order], bad])
    return offsets - base, out


def pad(starts, stops, content, length, clip=False):
    """jagged.py:1851-1900 -- (offsets, gathered content, invalid mask) of ``a.pad(length, clip=clip)``.

    The reference sizes every event to ``max(count, length)`` (or exactly ``length`` when clipping), builds
    parents / localindex of that layout, gathers ``content[starts[parents] + localindex]`` and masks the slots
    with ``localindex >= counts[parents]``.  What the gather reads under the mask is not part of the value; it is
    returned as 0 here.
    """
    starts, stops = np.asarray(starts, dtype=INDEX), np.asarray(stops, dtype=INDEX)
    content = np.asarray(content)
    counts = stops - starts
    sized = np.full(len(counts), length, dtype=INDEX) if clip else np.maximum(counts, length)
    offsets = counts2offsets(sized)
    parents = offsets2parents(offsets)
    local = np.arange(len(parents), dtype=INDEX) - offsets[:-1][parents]
    invalid = local >= counts[parents]
    index = starts[parents] + local
    data = np.zeros(len(parents), dtype=content.dtype)
    data[~invalid] = content[index[~invalid]]
    return offsets, data, invalid


def pad_fillna(starts, stops, content, length, value, clip=False):
    """``a.pad(length, clip=clip).fillna(value)`` (masked.py:401-406 with base.py:645-654): the padding AND any NaN
    of the content take ``value``, cast to the content dtype."""
    offsets, data, invalid = pad(starts, stops, content, length, clip)
    if len(np.asarray(content)) == 0:
        # jagged.py:1860-1864 wraps EMPTY content in an IndexedMaskedArray of -1s, whose fillna goes through
        # numpy.append(content, value) (masked.py:910-920): the dtype is promoted with the value's
        data = data.astype(np.result_type(data.dtype, np.asarray(value).dtype))
    if data.dtype.kind == "f":
        data[np.isnan(data)] = value
    data[invalid] = value
    return offsets, data


def parents2startsstops(parents):
    """jagged.py:73-93 -- children contiguous but not necessarily ordered or covering; negative parents skipped."""
    parents = np.asarray(parents, dtype=INDEX)
    tmp = np.nonzero(parents[1:] != parents[:-1])[0] + 1
    changes = np.empty(len(tmp) + 2, dtype=INDEX)
    changes[0], changes[-1], changes[1:-1] = 0, len(parents), tmp
    length = parents.max() + 1 if len(parents) > 0 else 0
    starts, counts = np.zeros(length, dtype=INDEX), np.zeros(length, dtype=INDEX)
    if len(parents) == 0:
        return starts, starts + counts
    where = parents[changes[:-1]]
    real = where >= 0
    starts[where[real]] = changes[:-1][real]
    counts[where[real]] = (changes[1:] - changes[:-1])[real]
    return starts, starts + counts


def uniques2offsetsparents(uniques):
    """jagged.py:95-110 -- runs of equal values are the events (no empty events)."""
    uniques = np.asarray(uniques)
    changes = np.nonzero(uniques[1:] != uniques[:-1])[0] + 1
    offsets = np.empty(len(changes) + 2, dtype=INDEX)
    offsets[0], offsets[-1], offsets[1:-1] = 0, len(uniques), changes
    parents = np.zeros(len(uniques), dtype=INDEX)
    parents[changes] = 1
    np.cumsum(parents, out=parents)
    return offsets, parents


def localindex2offsets(index):
    """jagged.py:194-222 (fromlocalindex): an event starts wherever the local index is 0."""
    index = np.asarray(index)
    if not ((index[1:] - index[:-1])[(index != 0)[1:]] == 1).all():
        raise ValueError("every localindex that is not zero must be one greater than the previous")
    starts = np.nonzero(index == 0)[0]
    offsets = np.empty(len(starts) + 1, dtype=INDEX)
    offsets[:-1], offsets[-1] = starts, len(index)
    return offsets


def folding2offsets(length, size):
    """jagged.py:237-242 (fromfolding)."""
    quotient = -(-length // size)
    offsets = np.arange(0, quotient * size + 1, size, dtype=INDEX)
    if len(offsets) > 0:
        offsets[-1] = length
    return offsets


def setitem(starts, stops, content, where_offsets, where, what):
    """jagged.py:800-833 -- ``a[where] = what`` for a jagged boolean mask or jagged integer index; returns the new
    content buffer (the reference assigns in place).  ``what`` is a scalar or one value per selected element."""
    starts, stops = np.asarray(starts, dtype=INDEX), np.asarray(stops, dtype=INDEX)
    content = np.array(content, copy=True)
    where_offsets, where = np.asarray(where_offsets, dtype=INDEX), np.asarray(where)
    counts = stops - starts
    wcounts = where_offsets[1:] - where_offsets[:-1]
    wparents = np.repeat(np.arange(len(wcounts), dtype=INDEX), wcounts)
    w = where[where_offsets[0]:where_offsets[-1]]
    if w.dtype == np.dtype(bool):
        if not np.array_equal(wcounts, counts):
            raise IndexError("jagged index does not fit")
        local = np.arange(len(w), dtype=INDEX) - (where_offsets[:-1] - where_offsets[0])[wparents]
        flat = (starts[wparents] + local)[w]
    else:
        idx = w.astype(INDEX)
        c = counts[wparents]
        idx = np.where(idx < 0, idx + c, idx)
        if not ((0 <= idx) & (idx < c)).all():
            raise IndexError("jagged array used as index contains out-of-bounds values")
        flat = idx + starts[wparents]
    content[flat] = what
    return content


def tolist(offsets, content):
    """base.py:159-172 -- nested python lists of a dense jagged array."""
    offsets = np.asarray(offsets)
    content = np.asarray(content)
    return [content[offsets[i]:offsets[i + 1]].tolist() for i in range(len(offsets) - 1)]


def synthetic(nevents, mean=5.0, seed=12345, dtype=np.float32, scale=25.0):
    """seeded synthetic input shared by the oracle and the CUDA path."""
    rng = np.random.default_rng(seed)
    counts = rng.poisson(mean, nevents).astype(INDEX)
    total = int(counts.sum())
    pt = rng.exponential(scale, total).astype(np.float32).astype(dtype)
    flat = rng.normal(size=nevents).astype(np.float32).astype(dtype)
    return counts, pt, flat
