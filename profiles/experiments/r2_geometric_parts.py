"""Where the geometric-tail rows lose their time: the same geometric counts (mean 4) without the 1000 long events,
with them, and with the long events alone appended at the end -- sum, argmax, a[a > 20], parents through the API."""
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
g.manual_seed(99)
n = 150_000_000


def run(tag, c):
    N = int(c.numel())
    M = int(c.sum())
    pt = torch.empty(M, device=dev, dtype=torch.float32).exponential_(1.0 / 25.0, generator=g)
    a = ak.JaggedArray.fromcounts(ak.DeviceArray(c), ak.DeviceArray(pt))
    off = a.offsets
    K = len(a[a > 20.0].content)
    npos = int((c > 0).sum())
    rows = (("sum", 4 * M + 8 * N + 4 * N, lambda: a.sum()),
            ("argmax", 4 * M + 8 * N + 8 * N + 8 * npos, lambda: a.argmax()),
            ("a[a>20]", 4 * M + 8 * N + 4 * K + 16 * N, lambda: a[a > 20.0]),
            ("parents", 8 * N + 8 * M, lambda: ak.JaggedArray.offsets2parents(off)))
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
        print("%-22s %-10s N=%d M=%d ms=%.4f frac=%.3f" % (tag, name, N, M, ms, nbytes / ms / 1e6 / 6552.3))
    del a, off, pt
    torch.cuda.empty_cache()


c = (torch.empty(n, device=dev, dtype=torch.float32).geometric_(0.2, generator=g).to(torch.int64) - 1).clamp_(min=0)
run("geometric", c)
c2 = c.clone()
where = torch.randint(0, n, (1000,), device=dev, generator=g)
longs = torch.randint(5000, 70000, (1000,), device=dev, generator=g)
c2[where] = longs
run("geometric+tail", c2)
c3 = c.clone()
where = torch.randint(0, n, (1000,), device=dev, generator=g)
c3[where] = 200
run("geometric+1000x200", c3)
