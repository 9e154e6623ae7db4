"""-m gpu: Parquet files (awkward0.toparquet / awkward0.fromparquet, arrow.py:452-578) to and from device arrays.
The golden files were written by the UNMODIFIED reference (oracle/make_golden_parquet.py); pq_expected.json is what
the reference's own reader returns for them."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

EXPECTED = json.load(open(os.path.join(GOLDEN_DIR, "pq_expected.json")))


@pytest.fixture(scope="module")
def ak():
    import awkward0_b200
    return awkward0_b200


def listify(x):
    """Rows of a Table come back as tuples of (name-ordered) values here and as dicts from the reference."""
    out = x.tolist()

    def fix(rows, names):
        if isinstance(rows, list):
            return [fix(r, names) for r in rows]
        if isinstance(rows, tuple):
            return dict(zip(names, rows))
        return rows
    names = x.columns
    return fix(out, names) if names else out


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_read_reference_files(ak, name):
    got = ak.fromparquet(os.path.join(GOLDEN_DIR, name + ".parquet"))
    assert isinstance(got, ak.JaggedArray)
    assert listify(got) == EXPECTED[name]


def test_same_file_layout_as_the_reference(ak, tmp_path):
    """What toparquet writes has the schema of the reference's file: ONE column named "" holding the lists."""
    import pyarrow.parquet as pq
    for name in ("pq_f64", "pq_i32", "pq_records", "pq_nested"):
        ref = os.path.join(GOLDEN_DIR, name + ".parquet")
        mine = str(tmp_path / (name + ".parquet"))
        ak.toparquet(mine, ak.fromparquet(ref))
        a, b = pq.read_table(ref), pq.read_table(mine)
        assert a.schema.names == b.schema.names == [""]
        assert a.column(0).to_pylist() == b.column(0).to_pylist()
        assert str(a.schema.field(0).type).replace("item", "element") == str(b.schema.field(0).type).replace("item", "element")


def test_round_trip_with_row_groups(ak, tmp_path):
    rng = np.random.default_rng(11)
    counts = rng.poisson(5.0, 200000).astype(np.int64)
    m = int(counts.sum())
    pt = ak.JaggedArray.fromcounts(counts, rng.exponential(25.0, m).astype(np.float32))
    sel = pt[pt > 20]                                   # a device result, never on the host before
    path = str(tmp_path / "sel.parquet")
    ak.toparquet(path, [sel[:70000], sel[70000:70001], sel[70001:]])
    import pyarrow.parquet as pq
    assert pq.ParquetFile(path).num_row_groups == 3
    back = ak.fromparquet(path)
    assert np.array_equal(back.counts.get(), sel.counts.get())
    assert np.array_equal(back.flatten().get(), sel.flatten().get())
    assert back.content.dtype == np.dtype(np.float32)
    assert np.allclose(back.sum().get(), sel.sum().get(), rtol=1e-5)
    with pytest.raises(ValueError):
        ak.toparquet(path, [])
    with pytest.raises(TypeError):
        ak.toparquet(path, 3)
    with pytest.raises(TypeError):
        ak.toparquet(path, [np.arange(3)])
