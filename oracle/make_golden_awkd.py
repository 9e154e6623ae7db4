#!/usr/bin/env python
"""Write tests/golden/*.awkd with the UNMODIFIED reference serializer  --  test infrastructure, build container only.

``awkward0.save`` itself cannot run under numpy 2 (persist.py:697 calls ``ndarray.tostring``, removed in numpy 2),
but it is a thin zip wrapper around ``awkward0.persist.serialize(array, storage, ...)`` (persist.py:497-503), which
takes any mapping as storage.  This script calls that serializer with the arguments ``save`` passes
(persist.py:688-690: delimiter "-", suffix ".raw", schemasuffix ".json", the default compression policy) and a
storage adapter that differs from ``save``'s in ONE call: ``ndarray.tobytes()`` (the same bytes) instead of
``tostring()``.  The files are then read back with the reference's own ``awkward0.load`` and the ``tolist()`` of
every member is written to tests/golden/awkd_expected.json.

    python oracle/make_golden_awkd.py
"""
import json
import os
import sys
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import _refloader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class Wrap(object):                      # persist.py:692-698 with tobytes() for tostring()
    def __init__(self, f):
        self.f = f

    def __setitem__(self, where, what):
        if isinstance(what, np.ndarray):
            what = what.tobytes()
        self.f.writestr(where, what, compress_type=zipfile.ZIP_STORED)


def save(ak0, path, arrays):
    import awkward0.persist as P
    if not isinstance(arrays, dict):
        arrays = {"": arrays}
    if os.path.exists(path):
        os.remove(path)
    with zipfile.ZipFile(path, mode="w", compression=zipfile.ZIP_STORED) as f:
        for name, array in arrays.items():
            P.serialize(array, Wrap(f), name=name, delimiter="-", suffix=".raw", schemasuffix=".json", compression=P.compression)


def main():
    ak0 = _refloader.load()
    J = ak0.JaggedArray
    rng = np.random.default_rng(2024)
    cases = {}
    cases["awkd_small_f32"] = J.fromcounts(np.array([3, 0, 2, 1], dtype=np.int64), np.array([1.5, 2.5, 3.5, 4.5, 5.5, 6.5], dtype=np.float32))
    counts = rng.poisson(3.0, 3000).astype(np.int64)                     # 24 kB of int64 counts: deflated
    m = int(counts.sum())
    cases["awkd_deflated_counts_f64"] = J.fromcounts(counts, rng.normal(size=m))
    cases["awkd_deflated_int_content"] = J.fromcounts(counts, rng.integers(-1000, 1000, m).astype(np.int32))
    c2 = rng.poisson(2.0, 400).astype(np.int64)
    m2 = int(c2.sum())
    pt = J.fromcounts(c2, rng.exponential(20.0, m2).astype(np.float32))
    cases["awkd_records"] = J.zip(pt=pt, eta=J.fromoffsets(pt.offsets, rng.uniform(-2.4, 2.4, m2).astype(np.float32)),
                                  q=J.fromoffsets(pt.offsets, rng.integers(0, 2, m2).astype(np.int8) * 2 - 1))
    inner = J.fromcounts(np.array([2, 0, 1, 3, 1], dtype=np.int64), np.arange(7, dtype=np.float64) * 1.5)
    cases["awkd_nested"] = J.fromcounts(np.array([2, 0, 3], dtype=np.int64), inner)
    cases["awkd_general_startsstops"] = J(np.array([4, 0, 2], dtype=np.int64), np.array([6, 2, 2], dtype=np.int64),
                                          np.array([0.0, 1.1, 2.2, 3.3, 4.4, 5.5, 6.6]))
    cases["awkd_same_small_lengths"] = J.fromcounts(np.array([1, 1, 1], dtype=np.int64), np.array([7, 8, 9], dtype=np.int64))
    cases["awkd_empty"] = J.fromcounts(np.array([], dtype=np.int64), np.array([], dtype=np.float32))
    cases["awkd_named"] = {"muons": cases["awkd_small_f32"], "hits": cases["awkd_same_small_lengths"]}
    expected = {}
    for name, arrays in cases.items():
        path = os.path.join(GOLDEN, name + ".awkd")
        save(ak0, path, arrays)
        loaded = ak0.load(path)                                      # the reference's own reader
        if isinstance(arrays, dict):
            expected[name] = dict((k, loaded[k].tolist()) for k in loaded)
            loaded.close()
        else:
            expected[name] = loaded.tolist()
        print(name, os.path.getsize(path), "bytes")
    with open(os.path.join(GOLDEN, "awkd_expected.json"), "w") as f:
        json.dump(expected, f)


if __name__ == "__main__":
    main()
