"""A slice of bench.py's per-kernel table on a device-drawn shard: `python profiles/experiments/r2_optable.py REGEX
[--events N] [--dist]` -- for kernel experiments that need three rows, not the whole bench."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "awkward-0.x_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
import awkward0_b200 as ak  # noqa: E402
from awkward0_b200 import _lib  # noqa: E402
from awkward0_b200.device import DeviceArray, empty, runtime  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("only")
ap.add_argument("--events", type=int, default=125_000_000)
ap.add_argument("--comb-events", type=int, default=50_000_000)
ap.add_argument("--kernel-reps", type=int, default=5)
ap.add_argument("--dist", action="store_true")
ap.add_argument("--dist-elements", type=int, default=600_000_000)
args = ap.parse_args()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(12345)
counts_t = torch.poisson(torch.full((args.events,), 5.0, device=dev), generator=g).to(torch.int64)
M = int(counts_t.sum())
pt_t = torch.empty(M, device=dev, dtype=torch.float32).exponential_(1.0 / 25.0, generator=g)
table = bench.kernel_table(ak, _lib, DeviceArray, empty, runtime, ak.DeviceArray(counts_t), ak.DeviceArray(pt_t), args.events, M, args)
if args.dist:
    table.update(bench.distribution_table(ak, args))
for k, v in table.items():
    print("%-42s ms=%.4f best=%.4f GB=%.3f frac=%.3f" % (k, v["ms"], v["best_ms"], v["bytes"] / 1e9, v["frac"]))
