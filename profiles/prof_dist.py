#!/usr/bin/env python
"""The four hot-path operations on bench.py's heavy-tailed shard (geometric counts, mean 4, plus 1000 events of 5 k - 70 k
elements; `python profiles/prof_dist.py poisson40` / `huge` for the other two): the command ncu lists launches of.

    python profiles/prof_dist.py [geometric_tail|poisson40|huge] [elements]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "awkward-0.x_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import awkward0_b200 as ak  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "geometric_tail"
elements = int(sys.argv[2]) if len(sys.argv) > 2 else 600_000_000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(99)
if which == "poisson40":
    c = torch.poisson(torch.full((elements // 40,), 40.0, device=dev), generator=g).to(torch.int64)
elif which == "huge":
    c = torch.full((1000,), max(1, elements // 1000), device=dev, dtype=torch.int64)
else:
    n = elements // 4
    c = (torch.empty(n, device=dev, dtype=torch.float32).geometric_(0.2, generator=g).to(torch.int64) - 1).clamp_(min=0)
    where = torch.randint(0, n, (1000,), device=dev, generator=g)
    c[where] = torch.randint(5000, 70000, (1000,), device=dev, generator=g)
M = int(c.sum())
x = torch.empty(M, device=dev, dtype=torch.float32).exponential_(1.0 / 25.0, generator=g)
a = ak.JaggedArray.fromcounts(ak.DeviceArray(c), ak.DeviceArray(x))
off = a.offsets
for name, fn in (("sum", lambda: a.sum()), ("argmax", lambda: a.argmax()), ("a[a > 20]", lambda: a[a > 20.0]),
                 ("parents", lambda: ak.JaggedArray.offsets2parents(off))):
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        del r
    print("%-10s %.3f ms" % (name, e0.elapsed_time(e1)))
print("ok", which, len(c), M)
