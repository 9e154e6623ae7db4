"""Which variant wins where: bench.distribution_table with the element-tiled kernels forced on every distribution
(JaggedArray.LONG_EVENTS_MEAN = 0) against the default choice (mean count > 7)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "awkward-0.x_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
import awkward0_b200 as ak  # noqa: E402


class Args:
    dist_elements = 600_000_000
    kernel_reps = 5


torch.cuda.set_device(0)
for thr in (7, 0):
    ak.JaggedArray.LONG_EVENTS_MEAN = thr
    table = bench.distribution_table(ak, Args)
    for k, v in table.items():
        print("mean>%d  %-42s ms=%.4f frac=%.3f max_count=%d" % (thr, k, v["ms"], v["frac"], v["max_count"]))
