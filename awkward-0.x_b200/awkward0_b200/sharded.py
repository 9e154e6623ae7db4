"""Event-range sharding of the JaggedArray hot path across GPUs (one process per GPU, torch.distributed).

The path shards by independent units: no operation moves an element across an event boundary, which is the
reference's own ChunkedArray model (awkward0/array/chunked.py:523-593, 602-658, 683-710 -- every op mapped
chunk by chunk; tests/test_crosscut.py:86-89) and why counts are position independent
(docs/classes.adoc:33-36).  Rank r owns a contiguous range of events with LOCAL offsets that start at 0; all
kernels run locally with no data-path collective.  NCCL is used only for
  * whole-array totals (sum of sums, global max/min, number selected, number of pairs): all_reduce of a few bytes;
  * replicated per-event results on request: all_gather (sizes first when shards are unequal); variable-size
    jagged results the same way, counts and content separately (`gather_jagged`).

Everything here works on whatever backend the process group uses: "nccl" with CUDA tensors in production,
"gloo" with CPU tensors in the world_size-2 CPU tests (tests/test_sharded_gloo.py), where the local compute
is supplied by the test.
"""
import numpy

try:
    import torch
    import torch.distributed as dist
except ImportError:          # pragma: no cover
    torch = None
    dist = None


def event_range(nevents, rank, world):
    """Contiguous, balanced event range [lo, hi) of `rank` (first nevents % world ranks get one more)."""
    if not 0 <= rank < world:
        raise ValueError("rank {0} outside world of size {1}".format(rank, world))
    base, extra = divmod(int(nevents), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def element_range(global_offsets, lo, hi):
    """Content range [offsets[lo], offsets[hi]) of an event range (host-side, on the global offsets)."""
    return int(global_offsets[lo]), int(global_offsets[hi])


def local_shard(counts, content, rank, world, offsets=None):
    """Slice host arrays (global counts, global dense content) into the shard of `rank`: (counts, content)."""
    counts = numpy.asarray(counts)
    lo, hi = event_range(len(counts), rank, world)
    if offsets is None:
        start = int(counts[:lo].sum())
        stop = start + int(counts[lo:hi].sum())
    else:
        start, stop = element_range(offsets, lo, hi)
    return counts[lo:hi], numpy.asarray(content)[start:stop]


def _world(group=None):
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def _device_for(group=None):
    if dist is not None and dist.is_initialized() and dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _as_tensor(x, dtype, device):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=dtype).reshape(-1)
    if hasattr(x, "tensor"):                   # DeviceArray
        return x.tensor.to(device=device, dtype=dtype).reshape(-1)
    return torch.as_tensor(numpy.asarray(x), dtype=dtype, device=device).reshape(-1)


def allreduce_totals(sums=(), maxima=(), minima=(), group=None):
    """Whole-array totals across ranks: returns (sums, maxima, minima) as python floats / ints.

    `sums` are added (sum of per-event sums, number selected, number of pairs ...), `maxima` / `minima`
    are reduced with max / min.  Payloads are a few bytes: the cost is launch latency, not bandwidth.
    """
    world, _ = _world(group)
    if world == 1:
        return tuple([v.item() if hasattr(v, "item") else v for v in vals] for vals in (sums, maxima, minima))
    device = _device_for(group)
    out = []
    for values, op in ((sums, "sum"), (maxima, "max"), (minima, "min")):
        values = list(values)
        if not values:
            out.append([])
            continue
        isint = all(isinstance(v, (int, numpy.integer)) for v in values)
        t = torch.tensor(values, dtype=torch.int64 if isint else torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op],
                            group=group)
        out.append([v.item() for v in t.cpu()])
    return tuple(out)


def gather_per_event(local, group=None, sizes=None):
    """Replicate a per-event result (one value per local event) on every rank, in global event order.

    ONE `all_gather_into_tensor` into a (world, biggest) buffer -- NCCL writes every rank's part straight into its
    row, no list of per-rank tensors and no staging copy when the shards are equal (the usual case: the row view IS
    the result).  Unequal shards are padded to the biggest and the rows trimmed afterwards.  `sizes` (events per
    rank) saves the 8-byte size exchange when the caller knows them (`event_range`).  `local` is a DeviceArray /
    torch tensor / numpy array; the result is a torch tensor on the group's device.
    """
    world, rank = _world(group)
    device = _device_for(group)
    if torch.is_tensor(local):
        t = local.to(device).reshape(-1)
    elif hasattr(local, "tensor"):
        t = local.tensor.to(device).reshape(-1)
    else:
        t = torch.as_tensor(numpy.asarray(local), device=device).reshape(-1)
    if world == 1:
        return t
    if sizes is None:
        mine = torch.tensor([t.numel()], dtype=torch.int64, device=device)
        every = torch.empty(world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(every, mine, group=group)
        sizes = [int(x) for x in every.cpu().tolist()]
    else:
        sizes = [int(x) for x in sizes]
        if len(sizes) != world or sizes[rank] != t.numel():
            raise ValueError("sizes must list the events of every rank")
    biggest = max(sizes)
    flat = torch.empty(world * biggest, dtype=t.dtype, device=device)      # (gloo insists on a flat output)
    out = flat.view(world, biggest)
    if t.numel() == biggest:
        src = t.contiguous()
    else:
        src = torch.zeros(biggest, dtype=t.dtype, device=device)
        src[:t.numel()] = t
    dist.all_gather_into_tensor(flat, src, group=group)
    if min(sizes) == biggest:
        return flat
    return torch.cat([out[r, :n] for r, n in enumerate(sizes)])


_PINNED = {}


def _to_host(flat):
    """A small uint8 tensor as a numpy array: from the device through a pinned buffer and an asynchronous copy (a
    pageable `.cpu()` costs ~0.1 ms per call on the bench host), from the host as it is (the gloo tests)."""
    if not flat.is_cuda:
        return flat.numpy()
    n = int(flat.numel())
    buf = _PINNED.get(flat.device.index)
    if buf is None or buf.numel() < n:
        buf = _PINNED[flat.device.index] = torch.empty(max(n, 4096), dtype=torch.uint8).pin_memory()
    buf[:n].copy_(flat, non_blocking=True)
    torch.cuda.current_stream(flat.device).synchronize()
    return buf[:n].numpy().copy()


class DeviceTotals(object):
    """Whole-array totals that never leave the device until the end of the step.

    Each partial (the sum of per-event sums, the number of selected elements, the number of leading objects ...) is
    already in HBM where a kernel epilogue wrote it -- a 1-element result of `jg_flat_reduce`, the last entry of a
    result's offsets (the number of selected elements, of leading objects ...).  `add_*` only remembers the device
    scalar; `finish()` packs them, exchanges the pack with ONE collective (`all_gather_into_tensor`: mixed integer /
    float slots and sum / max / min reductions ride in the same message as raw bytes) and reads it back with ONE
    device-to-host copy.  Replaces one D2H per partial plus an H2D, an all_reduce and a D2H per reduction kind
    (chunked.py:596-597 sums per-chunk results on the host the same way).
    """

    def __init__(self, group=None, capacity=None):
        self.group = group
        self.world, self.rank = _world(group)
        self.device = _device_for(group)
        self.parts = []                          # (tensor on the device, kind, numpy dtype), in slot order

    def _add(self, value, kind):
        if hasattr(value, "tensor"):
            value = value.tensor
        if torch.is_tensor(value):
            t = value.reshape(-1)[:1]
            if self.world > 1 and t.device != self.device:
                t = t.to(self.device)
            if t.dtype == torch.bool:
                t = t.to(torch.int64)
        else:                                    # a host number (sizes the host already had to know)
            dt = torch.int64 if isinstance(value, (int, numpy.integer)) else torch.float64
            t = torch.tensor([value], dtype=dt)
            if self.device.type != "cpu":
                t = t.to(self.device, non_blocking=True)
        self.parts.append((t, kind, numpy.dtype(str(t.dtype).replace("torch.", ""))))
        return len(self.parts) - 1

    def add_sum(self, value):
        return self._add(value, "sum")

    def add_max(self, value):
        return self._add(value, "max")

    def add_min(self, value):
        return self._add(value, "min")

    def finish(self):
        """-> list of python numbers, one per slot, reduced over all ranks: one packing launch, one all_gather, one copy."""
        if not self.parts:
            return []
        layout, at = [], 0
        for _, _, dt in self.parts:
            layout.append((at, dt))
            at += dt.itemsize
        # ONE launch packs every partial as raw bytes (no conversion: integers stay exact, a float32 stays float32)
        where = self.device
        if self.world == 1:                      # a single process packs where the partials already are (no copy)
            where = next((t.device for t, _, _ in self.parts if t.is_cuda), self.device)
        mine = torch.cat([t.to(where, non_blocking=True).view(torch.uint8) for t, _, _ in self.parts])
        nbytes = int(mine.numel())
        if self.world > 1:
            flat = torch.empty(self.world * nbytes, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(flat, mine.contiguous(), group=self.group)
        else:
            flat = mine
        host = _to_host(flat).reshape(self.world, nbytes)              # the step's ONE device-to-host copy
        out = []
        for j, (_, kind, _) in enumerate(self.parts):
            off, dt = layout[j]
            col = numpy.ascontiguousarray(host[:, off:off + dt.itemsize]).view(dt).reshape(-1)
            if kind == "sum":
                out.append(col.sum(dtype=numpy.float64 if dt.kind == "f" else numpy.int64).item())
            else:
                out.append((col.max() if kind == "max" else col.min()).item())
        return out


def gather_jagged(counts, content, group=None):
    """Replicate a VARIABLE-size jagged result (a selection, pairs, argmax ...) on every rank: the per-event counts and
    the dense content of every shard, each in global event order -- (counts, content) as torch tensors on the group's
    device; `fromcounts(counts, content)` is the whole-array result.  Two size exchanges + two padded all_gathers
    ("gather = allgather(sizes) + grouped send/recv"); results normally STAY sharded, this is on request."""
    return gather_per_event(counts, group=group), gather_per_event(content, group=group)


def global_offsets_base(local_total, group=None):
    """Exclusive prefix over ranks of the local element totals: what to add to local offsets to get global ones."""
    world, rank = _world(group)
    device = _device_for(group)
    if world == 1:
        return 0
    mine = torch.tensor([int(local_total)], dtype=torch.int64, device=device)
    every = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(every, mine, group=group)
    return int(sum(int(x.item()) for x in every[:rank]))


class ShardedJaggedArray(object):
    """This rank's shard of a jagged array plus the collectives that stitch whole-array answers together.

    `local` is any object with the JaggedArray reducer surface (the device JaggedArray in production).
    Per-event operations are simply forwarded to the shard (`s.local.sum()`, `s.local[s.local > 20]` ...).
    """

    def __init__(self, local, group=None):
        self.local = local
        self.group = group
        self.world, self.rank = _world(group)

    def __len__(self):
        (total,), _, _ = allreduce_totals(sums=[int(len(self.local))], group=self.group)
        return int(total)

    def _scalar(self, x):
        if hasattr(x, "get"):
            x = x.get()
        x = numpy.asarray(x).reshape(-1)
        return x[0].item() if len(x) else 0

    def total_sum(self):
        """Sum over all events of the per-event sums."""
        (tot,), _, _ = allreduce_totals(sums=[float(self._scalar(self.local.sum().sum()))], group=self.group)
        return tot

    def _extreme(self, ismin):
        """This shard's extreme as a python int / float; an EMPTY shard contributes the identity instead of raising, so
        that every rank always enters the collective (a rank that raised would leave the others waiting in NCCL)."""
        per_event = self.local.min() if ismin else self.local.max()       # identities for empty events
        if len(per_event) == 0:
            dt = numpy.dtype(getattr(per_event, "dtype", numpy.float64))
            if dt.kind in "iu":
                info = numpy.iinfo(dt)
                return int(info.max if ismin else info.min)
            return float("inf") if ismin else float("-inf")
        x = self._scalar(per_event.min() if ismin else per_event.max())
        return int(x) if isinstance(x, (int, numpy.integer)) else float(x)   # ints stay exact above 2^53

    def total_max(self):
        _, (m,), _ = allreduce_totals(maxima=[self._extreme(False)], group=self.group)
        return m

    def total_min(self):
        _, _, (m,) = allreduce_totals(minima=[self._extreme(True)], group=self.group)
        return m

    def total_content(self):
        """Number of content elements over all shards."""
        (tot,), _, _ = allreduce_totals(sums=[int(self._scalar(self.local.counts.sum()))], group=self.group)
        return int(tot)

    def gather(self, per_event):
        return gather_per_event(per_event, group=self.group)

    def gather_jagged(self, result):
        """A jagged per-shard result (`s.local[s.local > 20]`, `s.local.argmax()` ...) replicated whole on every rank,
        built with the result's own class (`fromcounts`)."""
        counts, content = gather_jagged(result.counts, result.flatten(), group=self.group)
        return type(result).fromcounts(counts, content)
