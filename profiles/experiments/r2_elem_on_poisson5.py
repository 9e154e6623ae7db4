"""The element-tiled variants on the bench's own Poisson(5) shard (forced with JaggedArray.LONG_EVENTS_MEAN = 0)
against the event-tiled ones: sum, argmax, a[a > 20], parents, localindex through the public API."""
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
g.manual_seed(12345)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000_000
mean = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
counts_t = torch.poisson(torch.full((N,), mean, device=dev), generator=g).to(torch.int64)
M = int(counts_t.sum())
pt_t = torch.empty(M, device=dev, dtype=torch.float32).exponential_(1.0 / 25.0, generator=g)
a = ak.JaggedArray.fromcounts(ak.DeviceArray(counts_t), ak.DeviceArray(pt_t))
off = a.offsets
K = len(a[a > 20.0].content)
npos = int((counts_t > 0).sum())
rows = (("sum", 4 * M + 8 * N + 4 * N, lambda: a.sum()),
        ("argmax", 4 * M + 8 * N + 8 * N + 8 * npos, lambda: a.argmax()),
        ("a[a>20]", 4 * M + 8 * N + 4 * K + 16 * N, lambda: a[a > 20.0]),
        ("parents", 8 * N + 8 * M, lambda: ak.JaggedArray.offsets2parents(off)),
        ("localindex", 8 * N + 8 * M, lambda: a.localindex))
for thr in (7, 0):
    ak.JaggedArray.LONG_EVENTS_MEAN = thr
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
        print("mean>%d  poisson(%g) %-12s ms=%.4f frac=%.3f" % (thr, mean, name, ms, nbytes / ms / 1e6 / 6552.3))
