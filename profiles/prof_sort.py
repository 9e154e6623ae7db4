#!/usr/bin/env python
"""Launch the 8f kernels (argsort, segment moves) on a 20 M-event shard: the command ncu profiles.

    python profiles/prof_sort.py [nevents]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "awkward-0.x_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import awkward0_b200 as ak  # noqa: E402
from awkward0_b200 import _lib  # noqa: E402
from awkward0_b200.device import empty, runtime  # noqa: E402
from awkward0_b200.jagged import _vp  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
rng = np.random.default_rng(1)
counts = rng.poisson(5.0, N).astype(np.int64)
M = int(counts.sum())
pt = (rng.standard_exponential(M, dtype=np.float32) * np.float32(25.0))
counts_d, pt_d = ak.asdevice(counts), ak.asdevice(pt)
a = ak.JaggedArray.fromcounts(counts_d, pt_d)
off = a._off64
st = runtime.stream
scal = runtime.scalars()
idx = empty(M, np.int64)
cat = empty(2 * M, np.float32)
dst = ak.DeviceArray(off.tensor[:-1] * 2)
for rep in range(3):
    _lib.call("jg_argsort", pt_d.data_ptr(), _lib.JG_F32, off.data_ptr(), N, 0, _vp(idx), _vp(scal, 1), _vp(scal, 2), 1, st())
    _lib.call("jg_segment_scatter", pt_d.data_ptr(), 4, off.data_ptr(), dst.data_ptr(), N, _vp(cat), st())
torch.cuda.synchronize()
print("ok", N, M)
