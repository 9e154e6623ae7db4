#!/usr/bin/env python
"""Launch the fused pair-expression kernel (jg_comb_eval: invariant mass on pairs of records) on a 20 M-event
shard: the command ncu profiles.

    python profiles/prof_eval.py [nevents]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "awkward-0.x_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import awkward0_b200 as ak  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
rng = np.random.default_rng(777)
mc = rng.poisson(2.0, N).astype(np.int64)
m = int(mc.sum())
first = ak.JaggedArray.fromcounts(ak.asdevice(mc), rng.exponential(20.0, m).astype(np.float32))
mk = lambda x: ak.JaggedArray.fromoffsets(first.offsets, x)
mu = ak.JaggedArray.zip(pt=first, eta=mk(rng.uniform(-2.4, 2.4, m).astype(np.float32)),
                        phi=mk(rng.uniform(-np.pi, np.pi, m).astype(np.float32)))


def mass(p):
    one, two = p.i0, p.i1
    return np.sqrt(2 * one.pt * two.pt * (np.cosh(one.eta - two.eta) - np.cos(one.phi - two.phi)))


for rep in range(3):
    p = mu.pairs()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mm = mass(p).content.tensor
    e1.record()
    torch.cuda.synchronize()
    print("fused mass: %.3f ms for %d pairs" % (e0.elapsed_time(e1), mm.numel()))
    del mm, p
print("ok", N, m)
