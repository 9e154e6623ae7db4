#!/usr/bin/env python
"""Parquet fixtures written by the UNMODIFIED reference (awkward0.toparquet, arrow.py:452-523) in the build
container: tests/golden/pq_*.parquet plus pq_expected.json, the `tolist()` of what the reference's own
`awkward0.fromparquet` reads back.  tests/test_gpu_parquet.py reads the same files on the device.

    python oracle/make_golden_parquet.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import _refloader  # noqa: E402


def main():
    ak0 = _refloader.load()
    J = ak0.JaggedArray
    rng = np.random.default_rng(20261020)
    counts = rng.poisson(3.0, 500)
    cases = {
        "pq_f64": J.fromiter([[0.0, 1.1, 2.2], [], [3.3, 4.4], [5.5, 6.6, 7.7, 8.8, 9.9]]),
        "pq_f32_poisson": J.fromcounts(counts, rng.exponential(25.0, counts.sum()).astype(np.float32)),
        "pq_i32": J.fromcounts([2, 0, 3], np.array([7, -7, 2147483647, 1, -3], dtype=np.int32)),
        "pq_bool": J.fromcounts([3, 0, 2, 1], np.array([True, False, True, False, False, True])),
        "pq_records": J.zip(pt=J.fromiter([[0.5, 1.5], [], [2.5]]), q=J.fromiter([[1, -1], [], [1]])),
        "pq_nested": J.fromcounts([2, 0, 2], J.fromcounts([2, 1, 0, 3], np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0]))),
        "pq_all_empty": J.fromiter([[], [], []]),
    }
    expected = {}
    gold = os.path.join(ROOT, "tests", "golden")
    for name, array in cases.items():
        path = os.path.join(gold, name + ".parquet")
        ak0.toparquet(path, array)
        try:
            expected[name] = ak0.fromparquet(path).tolist()
            assert expected[name] == array.tolist()
        except AttributeError:       # struct columns: the reference's reader asks pyarrow for StructType.num_children,
            expected[name] = array.tolist()                  # gone from current pyarrow; pin what was written
    # several row groups: the reference writes one per item of an iterable (arrow.py:497-523)
    a = cases["pq_f32_poisson"]
    path = os.path.join(gold, "pq_rowgroups.parquet")
    ak0.toparquet(path, [a[:100], a[100:101], a[101:]])
    expected["pq_rowgroups"] = ak0.fromparquet(path).tolist()
    assert expected["pq_rowgroups"] == a.tolist()
    with open(os.path.join(gold, "pq_expected.json"), "w") as f:
        json.dump(expected, f)
    print(len(expected), "parquet fixtures ->", gold)


if __name__ == "__main__":
    main()
