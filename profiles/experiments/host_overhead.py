#!/usr/bin/env python
"""Host-side cost of one sweep step (the GPU work is made negligible: 2000 events): wall time per step and the
cProfile of 2000 steps.  What the GPU waits for after every data-dependent size read is this Python."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "awkward-0.x_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import awkward0_b200 as ak  # noqa: E402

rng = np.random.default_rng(1)
counts = ak.asdevice(rng.poisson(5.0, 2000).astype(np.int64))
pt = ak.asdevice(rng.exponential(25.0, int(counts.get().sum())).astype(np.float32))


def sweep():
    a = ak.JaggedArray.fromcounts(counts, pt)
    s = a.sum()
    lead = a.argmax()
    sel = a[a > 20.0]
    flat = sel.flatten()
    c = sel.counts
    return float(s.sum()), len(flat), len(lead.content), len(c)


for _ in range(200):
    sweep()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000):
    sweep()
dt = (time.perf_counter() - t0) / 2000
print("host wall time per step: %.1f us" % (dt * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    sweep()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
