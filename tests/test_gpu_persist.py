"""-m gpu: `.awkd` files (awkward0.save / awkward0.load, persist.py:661-727) to and from device arrays.  The golden
files were written by the UNMODIFIED reference serializer; awkd_expected.json is what the reference's loader reads."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

EXPECTED = json.load(open(os.path.join(GOLDEN_DIR, "awkd_expected.json")))


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
def test_load_reference_files(ak, name):
    got = ak.load(os.path.join(GOLDEN_DIR, name + ".awkd"))
    want = EXPECTED[name]
    if isinstance(want, dict):
        assert sorted(got) == sorted(want)
        for k in want:
            assert listify(got[k]) == want[k]
        got.close()
    else:
        assert isinstance(got, ak.JaggedArray)
        assert listify(got) == want


def test_save_load_round_trip(ak, tmp_path):
    rng = np.random.default_rng(9)
    counts = rng.poisson(5.0, 200000).astype(np.int64)
    m = int(counts.sum())
    pt = ak.JaggedArray.fromcounts(counts, rng.exponential(25.0, m).astype(np.float32))
    mu = ak.JaggedArray.zip(pt=pt, charge=ak.JaggedArray.fromoffsets(pt.offsets, rng.integers(0, 2, m).astype(np.int8)))
    sel = pt[pt > 20]                                   # a device result, never on the host before
    path = str(tmp_path / "events")
    ak.save(path, {"mu": mu, "sel": sel, "flat": pt.sum()}, mode="w")
    assert os.path.exists(path + ".awkd")
    f = ak.load(path + ".awkd")
    assert sorted(f) == ["flat", "mu", "sel"]
    back = f["mu"]
    assert np.array_equal(back.offsets.get(), mu.offsets.get()) and back.columns == ["pt", "charge"]
    assert np.array_equal(back.pt.content.get(), mu.pt.content.get()) and back.charge.content.dtype == np.int8
    assert np.array_equal(f["sel"].content.get(), sel.content.get()) and np.array_equal(f["sel"].counts.get(), sel.counts.get())
    assert np.array_equal(f["flat"].get(), pt.sum().get())
    assert np.allclose(f["mu"].pt.sum().get(), pt.sum().get(), rtol=1e-5)      # and it is a working device array
    f.close()
    with pytest.raises(KeyError):
        ak.save(path, {"mu": mu})                       # appending an existing name is refused like the reference
    # general starts/stops survive as they are (jagged.py:296-302)
    g = pt[np.array([5, 1, 3])]
    ak.save(str(tmp_path / "general"), g, mode="w")
    h = ak.load(str(tmp_path / "general.awkd"))
    assert h.tolist() == g.tolist()
