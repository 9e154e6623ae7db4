"""Host side of one sweep step (round 2): (1) the step on a 2000-event shard -- kernels are negligible there, so its
time IS the host overhead + the four size read-backs; (2) cProfile of the same step on a 20 M-event shard."""
import cProfile, pstats, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awkward-0.x_b200"))
import numpy as np, torch
import awkward0_b200 as ak
from awkward0_b200 import sharded
import bench


def make(n):
    counts, pt = bench.make_shard(n, 1)
    return ak.asdevice(counts), ak.asdevice(pt)


def sweep(c, p):
    a = ak.JaggedArray.fromcounts(c, p)
    s = a.sum()
    acc = sharded.DeviceTotals()
    acc.add_sum(s.device_sum())
    sel = a[a > 20.0]
    lead = a.argmax()
    flat = sel.flatten(); cc = sel.counts
    acc.add_sum(sel.offsets[-1:]); acc.add_sum(lead.offsets[-1:])
    return acc.finish()


c, p = make(2000)
for _ in range(20): sweep(c, p)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): sweep(c, p)
torch.cuda.synchronize()
print("host overhead + 4 read-backs per step (2000-event shard): %.3f ms" % ((time.perf_counter() - t0) / 200 * 1e3))
c, p = make(20_000_000)
for _ in range(5): sweep(c, p)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(50): sweep(c, p)
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(30)
