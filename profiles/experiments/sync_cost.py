#!/usr/bin/env python
"""How much does reading one data-dependent size back cost?  (a) scal.cpu() as DeviceArray runtime.read does it,
(b) non_blocking copy into a pinned buffer + stream synchronize, (c) the kernel writes straight into pinned host
memory (UVA: cudaHostAlloc'ed memory is device-addressable) + stream synchronize."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "awkward-0.x_b200")):
    sys.path.insert(0, p)
import ctypes  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402
import awkward0_b200 as ak  # noqa: E402
from awkward0_b200 import _lib  # noqa: E402
from awkward0_b200.device import empty, runtime  # noqa: E402

n = 100000
counts = ak.asdevice(np.ones(n, dtype=np.int64))
out = empty(n + 1, np.int64)
ws, wsb = runtime.workspace(n)
dev_scal = torch.zeros(8, dtype=torch.int64, device="cuda")
pinned = torch.zeros(8, dtype=torch.int64).pin_memory()
stream = torch.cuda.current_stream()


def launch(total_ptr, status_ptr):
    _lib.call("jg_counts2offsets", counts.data_ptr(), _lib.JG_I64, n, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(total_ptr),
              ctypes.c_void_p(status_ptr), ws, wsb, runtime.stream())


def a():
    launch(dev_scal.data_ptr(), dev_scal.data_ptr() + 8)
    return int(dev_scal.cpu().numpy()[0])


def b():
    launch(dev_scal.data_ptr(), dev_scal.data_ptr() + 8)
    pinned.copy_(dev_scal, non_blocking=True)
    stream.synchronize()
    return int(pinned[0])


pinned_np = pinned.numpy()


def c():
    launch(pinned.data_ptr(), pinned.data_ptr() + 8)
    stream.synchronize()
    return int(pinned_np[0])


for name, fn in (("cpu()", a), ("pinned copy", b), ("mapped write", c), ("cpu()", a), ("mapped write", c)):
    assert fn() == n
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000):
        fn()
    dt = (time.perf_counter() - t0) / 2000
    print("%-14s %.1f us per launch + read" % (name, dt * 1e6))
