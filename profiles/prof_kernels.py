#!/usr/bin/env python
"""Launch every hot-path kernel a few times on a 20 M-event shard (Poisson(5), f32): the command ncu profiles.

    python profiles/prof_kernels.py [nevents]
"""
import ctypes
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
ws, wsb = runtime.workspace(N)
scal = runtime.scalars()
out_off = empty(N + 1, np.int64)
f32 = empty(N, np.float32)
i64a, i64b = empty(N + 1, np.int64), empty(N, np.int64)
mask = empty(M, np.bool_)
sel = empty(M, np.float32)
par = empty(M, np.int64)
for rep in range(3):
    _lib.call("jg_counts2offsets", counts_d.data_ptr(), _lib.JG_I64, N, _vp(out_off), _vp(scal), _vp(scal, 1), ws, wsb, st())
    _lib.call("jg_segreduce", pt_d.data_ptr(), _lib.JG_F32, off.data_ptr(), N, _lib.SUM, _vp(f32), M, st())
    _lib.call("jg_segreduce", pt_d.data_ptr(), _lib.JG_F32, off.data_ptr(), N, _lib.MAX, _vp(f32), M, st())
    _lib.call("jg_segargminmax_dense", pt_d.data_ptr(), _lib.JG_F32, off.data_ptr(), N, 0, _vp(i64a), _vp(i64b), ctypes.c_void_p(0), _vp(scal), _vp(scal, 1), ws, wsb, st())
    _lib.call("jg_binary", _lib.OP["GT"], pt_d.data_ptr(), _lib.JG_F32, ctypes.c_void_p(0), 0, _lib.RHS_SCALAR,
              ctypes.c_double(20.0), ctypes.c_int64(0), 0, _lib.JG_F32, _vp(mask), _lib.JG_B8, M, ctypes.c_void_p(0), 0, st())
    _lib.call("jg_mask_select", pt_d.data_ptr(), 4, _vp(mask), 0, off.data_ptr(), N, _vp(sel), _vp(i64a), ctypes.c_void_p(0), _vp(scal), ws, wsb, st())
    _lib.call("jg_offsets2counts", off.data_ptr(), N, _vp(i64b), st())
    _lib.call("jg_offsets2parents", off.data_ptr(), N, _vp(par), st())
    _lib.call("jg_broadcast", _vp(f32), 4, off.data_ptr(), N, _vp(sel), st())
# combinatorics on a dimuon-like shard (Poisson(2))
Nc = N // 2
mu_off = ak.JaggedArray.counts2offsets(ak.asdevice(rng.poisson(2.0, Nc).astype(np.int64)))
Mmu = int(mu_off[-1])
mu_pt = ak.asdevice(rng.exponential(20.0, Mmu).astype(np.float32))
poff = empty(Nc + 1, np.int64)
_lib.call("jg_comb_offsets", _lib.COMB_PAIRS, 2, mu_off.data_ptr(), ctypes.c_void_p(0), Nc, _vp(poff), _vp(scal), ws, wsb, st())
P = int(scal.cpu()[0])
c0, c1 = empty(P, np.int64), empty(P, np.int64)
arr = (ctypes.c_void_p * 2)(c0.data_ptr(), c1.data_ptr())
l, r = empty(P, np.float32), empty(P, np.float32)
flat = ak.asdevice(rng.normal(size=N).astype(np.float32))
lead = a.argmax()
gout = empty(len(lead.content), np.float32)
for rep in range(2):
    _lib.call("jg_comb_offsets", _lib.COMB_PAIRS, 2, mu_off.data_ptr(), ctypes.c_void_p(0), Nc, _vp(poff), _vp(scal), ws, wsb, st())
    _lib.call("jg_comb_fill", _lib.COMB_PAIRS, 2, mu_off.data_ptr(), ctypes.c_void_p(0), Nc, _vp(poff), arr, 0, st())
    _lib.call("jg_comb_gather2", _lib.COMB_PAIRS, mu_off.data_ptr(), ctypes.c_void_p(0), Nc, _vp(poff), mu_pt.data_ptr(), 4, mu_pt.data_ptr(), 4, _vp(l), _vp(r), st())
    _lib.call("jg_binary", _lib.OP["ADD"], pt_d.data_ptr(), _lib.JG_F32, flat.data_ptr(), _lib.JG_F32, _lib.RHS_PARENT,
              ctypes.c_double(0), ctypes.c_int64(0), 0, _lib.JG_F32, _vp(sel), _lib.JG_F32, M, off.data_ptr(), N, st())
    _lib.call("jg_index_gather", pt_d.data_ptr(), 4, off.data_ptr(), lead.content.data_ptr(), _lib.JG_I64,
              lead._off64.data_ptr(), N, _vp(gout), _vp(scal, 1), st())
    # the fused comparison + selection of the sweep (a[a > 20] as one op)
    _lib.call("jg_compare_select", pt_d.data_ptr(), 4, pt_d.data_ptr(), _lib.JG_F32, _lib.OP["GT"], ctypes.c_double(20.0), 0,
              off.data_ptr(), N, _vp(sel), _vp(i64a), _vp(i64b), _vp(scal), ws, wsb, st())
    _lib.call("jg_validate_offsets", off.data_ptr(), N, M, _vp(scal, 2), _vp(scal, 1), st())
# count distributions with long events: the element-tiled variants (round 2), through the public API
gen = torch.Generator(device="cuda")
gen.manual_seed(5)
c40 = torch.poisson(torch.full((N // 8,), 40.0, device="cuda"), generator=gen).to(torch.int64)
x40 = torch.empty(int(c40.sum()), device="cuda", dtype=torch.float32).exponential_(1.0 / 25.0, generator=gen)
b = ak.JaggedArray.fromcounts(ak.DeviceArray(c40), ak.DeviceArray(x40))
for rep in range(2):
    b.sum(); b.argmax(); b[b > 20.0]; ak.JaggedArray.offsets2parents(b.offsets)
torch.cuda.synchronize()
print("ok", N, M, P)
