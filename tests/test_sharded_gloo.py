"""not gpu: the N>1 host logic (event-range sharding + collectives) on a world_size-2 gloo group.

The local compute is the CPU oracle here (tests may use it); in production it is the CUDA JaggedArray and the
backend is NCCL.  The sharded answers must equal the single-process oracle on the concatenated input, which is
the reference's ChunkedArray semantics (tests/test_crosscut.py:86-89)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nevents, result_dir):
    for p in (ROOT, os.path.join(ROOT, "awkward-0.x_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from awkward0_b200 import sharded
    from oracle import jagged_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        counts, pt, _ = O.synthetic(nevents, seed=99)
        goff = O.counts2offsets(counts)
        lo, hi = sharded.event_range(nevents, rank, world)
        lcounts, lpt = sharded.local_shard(counts, pt, rank, world, offsets=goff)
        assert len(lcounts) == hi - lo and len(lpt) == goff[hi] - goff[lo]
        loff = O.counts2offsets(lcounts)                                  # LOCAL offsets start at 0
        lsum = O.reduce(loff, lpt, "sum")
        lmax = O.reduce(loff, lpt, "max")
        moff, mask = O.ufunc_broadcast(np.greater, loff, lpt, np.float32(20))
        soff, sel = O.mask_select(loff, lpt, moff, mask)
        local_max = float(lmax.max()) if len(lmax) else float("-inf")     # an empty shard contributes the identity
        (gsum, gk), (gmax,), _ = sharded.allreduce_totals(
            sums=[float(lsum.sum(dtype=np.float64)), int(len(sel))], maxima=[local_max])
        # the same totals through the one-collective / one-read-back buffer (device tensors and host numbers mixed)
        acc = sharded.DeviceTotals()
        acc.add_sum(torch.tensor([float(lsum.sum(dtype=np.float64))], dtype=torch.float64))
        acc.add_sum(int(len(sel)))
        acc.add_max(torch.tensor([local_max], dtype=torch.float32))
        acc.add_min(torch.tensor([2 ** 60 + 7 + rank], dtype=torch.int64))       # integers stay exact above 2^53
        dsum, dk, dmax, dmin = acc.finish()
        assert dsum == gsum and dk == gk and dmax == gmax and dmin == 2 ** 60 + 7 and isinstance(dk, int)
        lo_sizes = [b - a for a, b in (sharded.event_range(nevents, r, world) for r in range(world))]
        assert np.array_equal(sharded.gather_per_event(lsum, sizes=lo_sizes).numpy(), sharded.gather_per_event(lsum).numpy())
        every_sum = sharded.gather_per_event(lsum).numpy()
        sel_counts, sel_content = sharded.gather_jagged(O.offsets2counts(soff), sel)      # variable-size result
        base = sharded.global_offsets_base(len(lpt))
        assert base == goff[lo]

        class Local(object):            # the reducer surface ShardedJaggedArray forwards to
            counts = lcounts

            def __len__(self):
                return len(lcounts)

            def sum(self):
                return lsum

            def max(self):
                return lmax

            def min(self):
                return O.reduce(loff, lpt, "min")

        class Result(object):           # a jagged result with the surface gather_jagged uses
            def __init__(self, counts, content):
                self.counts, self.content = counts, content

            def flatten(self):
                return self.content

            @classmethod
            def fromcounts(cls, counts, content):
                return cls(counts, content)

        s = sharded.ShardedJaggedArray(Local())
        whole = s.gather_jagged(Result(O.offsets2counts(soff), sel))
        assert np.array_equal(whole.counts.numpy(), sel_counts.numpy()) and np.array_equal(whole.content.numpy(), sel_content.numpy())
        np.savez(os.path.join(result_dir, "rank%d.npz" % rank), gsum=gsum, gk=gk, gmax=gmax, every_sum=every_sum,
                 sel_counts=sel_counts.numpy(), sel_content=sel_content.numpy(),
                 n=len(s), tsum=s.total_sum(), tmax=s.total_max(), tmin=s.total_min(), tcontent=s.total_content())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nevents", [10001, 10000, 4, 1])      # unequal / equal shards, tiny, and an EMPTY shard on rank 1
def test_two_rank_gloo_matches_single_process(tmp_path, nevents):
    from oracle import jagged_oracle as O
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), nevents, str(tmp_path)), nprocs=world, join=True)
    counts, pt, _ = O.synthetic(nevents, seed=99)
    off = O.counts2offsets(counts)
    want_sum = O.reduce(off, pt, "sum")
    moff, mask = O.ufunc_broadcast(np.greater, off, pt, np.float32(20))
    soff, sel = O.mask_select(off, pt, moff, mask)
    for rank in range(world):
        r = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert np.isclose(float(r["gsum"]), float(want_sum.sum(dtype=np.float64)), rtol=1e-6)
        assert int(r["gk"]) == len(sel)
        assert float(r["gmax"]) == float(O.reduce(off, pt, "max").max())
        assert np.array_equal(r["every_sum"], want_sum)              # per-event sums: bit-exact, in global event order
        assert np.array_equal(r["sel_counts"], O.offsets2counts(soff)) and np.array_equal(r["sel_content"], sel)
        assert int(r["n"]) == nevents and int(r["tcontent"]) == len(pt)
        assert np.isclose(float(r["tsum"]), float(want_sum.sum(dtype=np.float64)), rtol=1e-6)
        assert float(r["tmax"]) == float(pt.max()) and float(r["tmin"]) == float(pt.min())


def test_event_range_partition():
    from awkward0_b200 import sharded
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            ranges = [sharded.event_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharded.event_range(10, 2, 2)
