"""a + flat (the broadcast is never materialised: binary_parent_kernel) at several mean counts, through the API."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "awkward-0.x_b200")]
import torch  # noqa: E402
import awkward0_b200 as ak  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(5)
for mean, N in ((5.0, 125_000_000), (10.0, 60_000_000), (20.0, 30_000_000), (40.0, 15_000_000)):
    c = torch.poisson(torch.full((N,), mean, device=dev), generator=g).to(torch.int64)
    M = int(c.sum())
    pt = torch.empty(M, device=dev, dtype=torch.float32).exponential_(1.0 / 25.0, generator=g)
    flat = ak.DeviceArray(torch.rand(N, device=dev, dtype=torch.float32, generator=g))
    a = ak.JaggedArray.fromcounts(ak.DeviceArray(c), ak.DeviceArray(pt))
    rows = (("a + flat", 8 * M + 4 * N + 8 * N, lambda: a + flat), ("a > flat", 5 * M + 4 * N + 8 * N, lambda: a > flat))
    for name, nbytes, fn in rows:
        r = fn()
        del r
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            del r
        ms = float(np.mean(ts))
        print("poisson(%g) %-9s ms=%.4f frac=%.3f" % (mean, name, ms, nbytes / ms / 1e6 / 6552.3))
    del a, pt, c, flat
    torch.cuda.empty_cache()
