#!/usr/bin/env python
"""Launch the combinatorics kernels (value pairs / distincts / cross, argpairs, argchoose(3)) on 20 M Poisson(2) events
(argchoose: 5 M Poisson(5) events): the command ncu profiles.

    python profiles/prof_comb.py [nevents]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "awkward-0.x_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import awkward0_b200 as ak  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
rng = np.random.default_rng(777)
mc = rng.poisson(2.0, N).astype(np.int64)
mu = ak.JaggedArray.fromcounts(ak.asdevice(mc), rng.exponential(20.0, int(mc.sum())).astype(np.float32))
ec = rng.poisson(2.0, N).astype(np.int64)
el = ak.JaggedArray.fromcounts(ak.asdevice(ec), rng.exponential(20.0, int(ec.sum())).astype(np.float32))
jc = rng.poisson(5.0, N // 4).astype(np.int64)
jets = ak.JaggedArray.fromcounts(ak.asdevice(jc), rng.exponential(20.0, int(jc.sum())).astype(np.float32))


def timed(name, fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-24s %.3f ms" % (name, e0.elapsed_time(e1)))
    return r


timed("pairs values", lambda: (lambda p: (p.i0.content, p.i1.content))(mu.pairs()))
timed("distincts values", lambda: (lambda p: (p.i0.content, p.i1.content))(mu.distincts()))
timed("cross values", lambda: (lambda p: (p.i0.content, p.i1.content))(mu.cross(el)))
timed("argpairs", lambda: mu.argpairs())
timed("argchoose(3)", lambda: jets.argchoose(3))
print("ok", N)
