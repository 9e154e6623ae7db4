"""cProfile of the Python host path of one sweep (where do the ~0.45 ms/step outside kernels go?)."""
import cProfile, pstats, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "awkward-0.x_b200"))
import numpy as np, torch
import awkward0_b200 as ak
from awkward0_b200 import sharded
import bench
counts, pt = bench.make_shard(20_000_000, 1)
c, p = ak.asdevice(counts), ak.asdevice(pt)
def sweep():
    a = ak.JaggedArray.fromcounts(c, p)
    s = a.sum(); lead = a.argmax(); sel = a[a > 20.0]; flat = sel.flatten(); cc = sel.counts
    return float(s.sum()), len(flat), len(lead.content), len(cc)
for _ in range(5): sweep()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(50): sweep()
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
