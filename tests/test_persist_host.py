"""CPU tests of the `.awkd` container code (awkward0_b200/persist.py, host half): the golden files were written by
the UNMODIFIED reference serializer (oracle/make_golden_awkd.py) and awkd_expected.json holds what the reference's
own loader makes of them.  When the reference tree is present (build container) files written here are also loaded
by the reference."""
import io
import json
import os
import zipfile

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from awkward0_b200 import persist as P

EXPECTED = json.load(open(os.path.join(GOLDEN_DIR, "awkd_expected.json")))


def tree_tolist(t):
    if isinstance(t, np.ndarray):
        return t.tolist()
    if isinstance(t, P.HostTable):
        cols = dict((n, tree_tolist(c)) for n, c in t.columns.items())
        n = len(next(iter(cols.values()))) if cols else 0
        return [dict((k, v[i]) for k, v in cols.items()) for i in range(n)]
    content = tree_tolist(t.content)
    if t.counts is not None:
        off = np.concatenate([[0], np.cumsum(t.counts)])
        return [content[off[i]:off[i + 1]] for i in range(len(t.counts))]
    return [content[a:b] for a, b in zip(t.starts.tolist(), t.stops.tolist())]


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_reads_what_the_reference_wrote(name):
    got = P.load_host(os.path.join(GOLDEN_DIR, name + ".awkd"))
    want = EXPECTED[name]
    if isinstance(want, dict):
        assert sorted(got) == sorted(want)
        for k in want:
            assert tree_tolist(got[k]) == want[k]
    else:
        assert tree_tolist(got) == want


def test_deflate_policy_and_layout_match_the_reference():
    with zipfile.ZipFile(os.path.join(GOLDEN_DIR, "awkd_deflated_counts_f64.awkd")) as f:
        ref_schema = json.loads(f.read(".json"))
        ref_names = sorted(f.namelist())
    tree = P.load_host(os.path.join(GOLDEN_DIR, "awkd_deflated_counts_f64.awkd"))
    buf = io.BytesIO()
    P.save_host(buf, {"": tree}, mode="w")
    with zipfile.ZipFile(buf) as f:
        mine = json.loads(f.read(".json"))
        assert sorted(f.namelist()) == ref_names
        assert all(i.compress_type == zipfile.ZIP_STORED for i in f.infolist())
    assert mine == ref_schema                      # same calls, same ids, counts deflated, float64 content raw
    assert tree_tolist(P.load_host(buf)) == EXPECTED["awkd_deflated_counts_f64"]


def test_round_trip_records_named_and_general():
    rng = np.random.default_rng(1)
    counts = rng.poisson(2.0, 5000).astype(np.int64)
    m = int(counts.sum())
    rec = P.HostJagged(P.HostTable([("pt", rng.normal(size=m).astype(np.float32)), ("q", rng.integers(-1, 2, m).astype(np.int8))]), counts=counts)
    gen = P.HostJagged(np.arange(10.0), starts=np.array([5, 0], dtype=np.int64), stops=np.array([9, 3], dtype=np.int64))
    buf = io.BytesIO()
    P.save_host(buf, {"rec": rec, "gen": gen}, mode="w")
    back = P.load_host(buf)
    assert list(back) == ["rec", "gen"]
    assert tree_tolist(back["rec"]) == tree_tolist(rec) and tree_tolist(back["gen"]) == [[5.0, 6.0, 7.0, 8.0], [0.0, 1.0, 2.0]]
    assert back["rec"].content.columns["q"].dtype == np.int8
    with pytest.raises(KeyError):
        P.save_host(buf, {"rec": rec}, mode="a")                       # persist.py:700-704: no silent overwrite
    with pytest.raises(KeyError):
        P.save_host(io.BytesIO(), {"a": gen, "ab": gen}, mode="w")     # persist.py:672-676: prefix clash


def test_only_hot_path_nodes_are_accepted():
    def storage(schema):
        return {".json": json.dumps({"awkward0": "0.15.5", "schema": schema}).encode("ascii"), "1.raw": b"\0" * 8}
    arr = {"call": ["awkward0", "numpy", "frombuffer"], "args": [{"read": "1.raw"}, {"dtype": "int64"}, {"json": 1}]}
    assert P.read_tree(storage(arr)).tolist() == [0]
    for bad in ({"call": ["awkward0", "ChunkedArray"], "args": [{"list": [arr]}]},
                {"call": ["awkward0", "VirtualArray"], "args": []},
                {"python": "gASVBQAAAAAAAABLAS4="},
                {"call": ["os", "system"], "args": [{"json": "true"}]},
                {"call": ["awkward0", "numpy", "frombuffer"], "args": [{"call": ["lzma", "decompress"], "args": [{"read": "1.raw"}]}, {"dtype": "int64"}, {"json": 1}]}):
        with pytest.raises(NotImplementedError):
            P.read_tree(storage(bad))
    with pytest.raises(ValueError):
        P.read_tree({".json": b'{"schema": {}}'})


def test_reference_loads_what_we_write(tmp_path):
    from oracle import _refloader
    if not _refloader.available():
        pytest.skip("reference tree not present (GPU box)")
    ak0 = _refloader.load()
    rng = np.random.default_rng(3)
    counts = rng.poisson(4.0, 4000).astype(np.int64)
    m = int(counts.sum())
    rec = P.HostJagged(P.HostTable([("pt", rng.normal(size=m)), ("n", rng.integers(0, 9, m).astype(np.int32))]), counts=counts)
    nested = P.HostJagged(P.HostJagged(np.arange(6.0), counts=np.array([1, 2, 0, 3], dtype=np.int64)), counts=np.array([3, 0, 1], dtype=np.int64))
    path = str(tmp_path / "mine.awkd")
    P.save_host(path, {"rec": rec, "nested": nested}, mode="w")
    f = ak0.load(path)
    assert f["nested"].tolist() == [[[0.0], [1.0, 2.0], []], [], [[3.0, 4.0, 5.0]]]
    assert f["rec"].tolist() == tree_tolist(rec)
    f.close()
