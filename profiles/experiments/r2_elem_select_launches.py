"""One a[a > 20] on Poisson(40) counts (15 M events, 600 M float32): run under
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum to list the launches of the element-tiled selection."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "awkward-0.x_b200")]
import torch  # noqa: E402
import awkward0_b200 as ak  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(99)
n = 15_000_000
c = torch.poisson(torch.full((n,), 40.0, device=dev), generator=g).to(torch.int64)
M = int(c.sum())
pt = torch.empty(M, device=dev, dtype=torch.float32).exponential_(1.0 / 25.0, generator=g)
a = ak.JaggedArray.fromcounts(ak.DeviceArray(c), ak.DeviceArray(pt))
for _ in range(3):
    r = a[a > 20.0]
    del r
    r = a.argmax()
    del r
torch.cuda.synchronize()
print("done", M)
