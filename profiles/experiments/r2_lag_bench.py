"""Round 2 experiment: lagged persistent scan / selection vs the round-1 look-back scan and count/scan/fill selection.

    JG_SCAN_MODE={0,1,2} JG_SELECT_MODE={0,1,2} [JG_SCAN_CTAS=k] [JG_SELECT_CTAS=k] python profiles/experiments/r2_lag_bench.py [N]

Prints one line per op: mean / best ms over 7 launches, GB/s of algorithmic bytes, fraction of the measured HBM peak,
and whether the result equals torch's (cumsum / boolean compaction) bit for bit.
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "awkward-0.x_b200"))
import awkward0_b200 as ak                                    # noqa: E402
from awkward0_b200 import _lib                                # noqa: E402
from awkward0_b200.device import empty, runtime              # noqa: E402
from awkward0_b200.jagged import _vp                          # noqa: E402

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 125_000_000
MEAN = float(os.environ.get("EXP_MEAN", "5"))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6552.3
tag = "scan=%s sel=%s sctas=%s selctas=%s" % (os.environ.get("JG_SCAN_MODE", "-"), os.environ.get("JG_SELECT_MODE", "-"),
                                              os.environ.get("JG_SCAN_CTAS", "-"), os.environ.get("JG_SELECT_CTAS", "-"))

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1234)
counts_t = torch.poisson(torch.full((N,), MEAN, device=dev), generator=g).to(torch.int64)
M = int(counts_t.sum())
pt_t = torch.empty(M, device=dev, dtype=torch.float32).exponential_(1.0 / 25.0, generator=g)
counts_dev, pt_dev = ak.DeviceArray(counts_t), ak.DeviceArray(pt_t)
st = runtime.stream
ws, wsb = runtime.workspace(N)
scal = runtime.scalars()


def time_call(fn, reps=7):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts)), float(min(ts))


def report(name, nbytes, fn, ok):
    ms, best = time_call(fn)
    print("%-28s %-34s %8.4f ms (best %8.4f) %7.0f GB/s  frac %.3f  %s" % (
        name, tag, ms, best, nbytes / ms / 1e6, nbytes / ms / 1e6 / PEAK, "OK" if ok else "MISMATCH"), flush=True)


# ---- counts2offsets ---------------------------------------------------------------------------------------------
off = empty(N + 1, np.int64)
call = lambda: _lib.call("jg_counts2offsets", counts_dev.data_ptr(), _lib.JG_I64, N, _vp(off), _vp(scal), _vp(scal, 1), ws, wsb, st())
call()
ref_off = torch.zeros(N + 1, dtype=torch.int64, device=dev)
torch.cumsum(counts_t, 0, out=ref_off[1:])
ok = bool(torch.equal(off, ref_off)) and int(scal.cpu()[0]) == M
report("counts2offsets", 16 * N, call, ok)

tiles = empty((N + 511) // 512, np.int64)
off2 = empty(N + 1, np.int64)
call = lambda: _lib.call("jg_counts2offsets_tiles", counts_dev.data_ptr(), _lib.JG_I64, N, _vp(off2), _vp(tiles), _vp(scal), _vp(scal, 1), ws, wsb, st())
call()
ne = (counts_t > 0).to(torch.int64)
pad = (-N) % 512
ne_tiles = torch.cat([ne, torch.zeros(pad, dtype=torch.int64, device=dev)]).view(-1, 512).sum(1)
ok = bool(torch.equal(off2, ref_off)) and bool(torch.equal(tiles, ne_tiles))
report("counts2offsets_tiles", 16 * N, call, ok)
del off2, ne, ne_tiles

c32 = ak.DeviceArray(counts_t.to(torch.int32))
off3 = empty(N + 1, np.int64)
call = lambda: _lib.call("jg_counts2offsets", c32.data_ptr(), _lib.JG_I32, N, _vp(off3), _vp(scal), _vp(scal, 1), ws, wsb, st())
call()
report("counts2offsets_i32", 12 * N, call, bool(torch.equal(off3, ref_off)))
del c32, off3

# ---- comb offsets -----------------------------------------------------------------------------------------------
Nc = min(N, 50_000_000)
ca = torch.poisson(torch.full((Nc,), 2.0, device=dev), generator=g).to(torch.int64)
cb = torch.poisson(torch.full((Nc,), 2.0, device=dev), generator=g).to(torch.int64)
oa = torch.zeros(Nc + 1, dtype=torch.int64, device=dev)
ob = torch.zeros(Nc + 1, dtype=torch.int64, device=dev)
torch.cumsum(ca, 0, out=oa[1:])
torch.cumsum(cb, 0, out=ob[1:])
oa_d, ob_d = ak.DeviceArray(oa), ak.DeviceArray(ob)
wsc, wscb = runtime.workspace(Nc)
for kind, name, second in ((_lib.COMB_PAIRS, "pairs", None), (_lib.COMB_CROSS, "cross", ob_d)):
    poff = empty(Nc + 1, np.int64)
    obp = second.data_ptr() if second is not None else ctypes.c_void_p(0)
    call = lambda: _lib.call("jg_comb_offsets", kind, 2, oa_d.data_ptr(), obp, Nc, _vp(poff), _vp(scal), wsc, wscb, st())
    call()
    per = ca * (ca + 1) // 2 if second is None else ca * cb
    ref = torch.zeros(Nc + 1, dtype=torch.int64, device=dev)
    torch.cumsum(per, 0, out=ref[1:])
    report("comb_offsets_" + name, (16 if second is None else 24) * Nc, call, bool(torch.equal(poff, ref)))
    del poff, ref, per
del ca, cb, oa, ob, oa_d, ob_d

# ---- selection ----------------------------------------------------------------------------------------------------
mask_t = pt_t > 20.0
K = int(mask_t.sum())
ref_sel = pt_t[mask_t]
seg = torch.repeat_interleave(torch.arange(N, device=dev), counts_t)
newc = torch.zeros(N, dtype=torch.int64, device=dev)
newc.index_add_(0, seg[mask_t], torch.ones(K, dtype=torch.int64, device=dev))
del seg
ref_noff = torch.zeros(N + 1, dtype=torch.int64, device=dev)
torch.cumsum(newc, 0, out=ref_noff[1:])
mask = ak.DeviceArray(mask_t)
offd = ak.DeviceArray(ref_off)
sel = empty(M, np.float32)
noff = empty(N + 1, np.int64)
ncnt = empty(N, np.int64)


def check_sel():
    return (int(scal.cpu()[0]) == K and bool(torch.equal(sel[:K], ref_sel)) and bool(torch.equal(noff, ref_noff))
            and bool(torch.equal(ncnt, newc)))


call = lambda: _lib.call("jg_mask_select", pt_dev.data_ptr(), 4, mask.data_ptr(), 0, offd.data_ptr(), N, _vp(sel), _vp(noff), _vp(ncnt), _vp(scal), ws, wsb, st())
call()
report("mask_select_f32", 4 * M + M + 8 * N + 4 * K + 16 * N, call, check_sel())
sel.zero_(); noff.zero_(); ncnt.zero_()
call = lambda: _lib.call("jg_compare_select", pt_dev.data_ptr(), 4, pt_dev.data_ptr(), _lib.JG_F32, _lib.OP["GT"], ctypes.c_double(20.0), 0,
                         offd.data_ptr(), N, _vp(sel), _vp(noff), _vp(ncnt), _vp(scal), ws, wsb, st())
call()
report("compare_select_f32", 4 * M + 8 * N + 4 * K + 16 * N, call, check_sel())
del sel, ref_sel
pt64 = ak.DeviceArray(pt_t.double())
sel64 = empty(M, np.float64)
call = lambda: _lib.call("jg_compare_select", pt64.data_ptr(), 8, pt64.data_ptr(), _lib.JG_F64, _lib.OP["GT"], ctypes.c_double(20.0), 0,
                         offd.data_ptr(), N, _vp(sel64), _vp(noff), _vp(ncnt), _vp(scal), ws, wsb, st())
call()
ok = int(scal.cpu()[0]) == K and bool(torch.equal(sel64[:K], pt64.tensor[mask_t])) and bool(torch.equal(noff, ref_noff))
report("compare_select_f64", 8 * M + 8 * N + 8 * K + 16 * N, call, ok)
call = lambda: _lib.call("jg_mask_select", pt64.data_ptr(), 8, mask.data_ptr(), 0, offd.data_ptr(), N, _vp(sel64), _vp(noff), _vp(ncnt), _vp(scal), ws, wsb, st())
call()
ok = int(scal.cpu()[0]) == K and bool(torch.equal(sel64[:K], pt64.tensor[mask_t])) and bool(torch.equal(noff, ref_noff))
report("mask_select_f64", 8 * M + M + 8 * N + 8 * K + 16 * N, call, ok)
