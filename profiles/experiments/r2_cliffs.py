"""Looking for cliffs: event-tiled operations that have no element-tiled variant, at mean counts of 5 / 10 / 20 / 40 --
ms per million content elements through the API (a flat profile means no cliff)."""
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
M0 = 200_000_000
for mean in (5.0, 10.0, 20.0, 40.0):
    N = int(M0 / mean)
    c = torch.poisson(torch.full((N,), mean, device=dev), generator=g).to(torch.int64)
    M = int(c.sum())
    pt = torch.empty(M, device=dev, dtype=torch.float32).exponential_(1.0 / 25.0, generator=g)
    a = ak.JaggedArray.fromcounts(ak.DeviceArray(c), ak.DeviceArray(pt))
    idx = a.localindex
    rev = a[::-1] if False else None
    b = ak.JaggedArray.fromcounts(ak.DeviceArray(c), ak.DeviceArray(pt * 2))
    rows = [("a[localindex]", lambda: a[idx]),
            ("argsort", lambda: a.argsort()),
            ("concat axis=1", lambda: ak.JaggedArray.concatenate([a, b], axis=1)),
            ("a + b (same counts)", lambda: a + b),
            ("a[1:] events view compact", lambda: a[1:].compact()),
            ("pad(12)", lambda: a.pad(12)),
            ("a * a.max()", lambda: a * a.max()),
            ("counts", lambda: a[a > 20.0].counts)]
    if mean <= 10:
        rows.append(("argpairs", lambda: a.argpairs()))
    for name, fn in rows:
        try:
            r = fn()
            del r
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
                del r
            print("poisson(%g) %-28s ms=%8.3f  ms per M elements=%.5f" % (mean, name, float(np.mean(ts)), float(np.mean(ts)) / (M / 1e6)))
        except Exception as e:
            print("poisson(%g) %-28s failed: %r" % (mean, name, e))
        torch.cuda.empty_cache()
    del a, b, idx, pt, c
    torch.cuda.empty_cache()
