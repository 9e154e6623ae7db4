"""Host -> HBM ingest of a (counts, content) pair with the PCIe copy overlapped with the kernels.

The hot path is event-separable (awkward0/array/chunked.py:523-593: every operation of the reference's ChunkedArray
is mapped chunk by chunk), so a jagged array that still lives in host memory does not have to arrive completely
before work starts.  `StreamedJaggedArray` queues ONE copy of `counts` and then the content in fixed-size pieces on a
side stream, scans the counts as soon as they have landed (jg_counts2offsets), and hands out the array as
consecutive event-range VIEWS -- same offsets buffer, same content buffer, offsets[0] != 0 -- each of which becomes
available when the piece holding its last element has been copied.  Nothing is copied twice and the views alias the
one resident array, which is also available whole (`.whole()`) once the last piece is in.

    for chunk in StreamedJaggedArray(counts_pinned, pt_pinned, nchunks=8):
        s = chunk.sum(); sel = chunk[chunk > 20]        # runs while later pieces are still on the bus

The copy engine and the SMs work concurrently, so the cost of a step from host memory drops from copy + compute to
about the copy alone.  Pinned host buffers (torch tensors with pin_memory=True) are needed for the overlap; pageable
numpy arrays work too but are staged by the driver.
"""
import numpy

from . import _lib
from .device import DeviceArray, INDEXTYPE, empty, runtime, torch, torch_dtype, _TORCH2NP
from .jagged import JaggedArray


def _as_host_tensor(x):
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            raise TypeError("StreamedJaggedArray takes HOST buffers; device arrays need no ingest")
        return x.reshape(-1)
    arr = numpy.ascontiguousarray(numpy.asarray(x))
    if arr.dtype.kind == "u" and arr.dtype.itemsize > 1:
        return torch.from_numpy(arr.view("i%d" % arr.dtype.itemsize)).view(torch_dtype(arr.dtype))
    return torch.from_numpy(arr)


class StreamedJaggedArray(object):
    def __init__(self, counts, content, nchunks=8, piece_bytes=64 << 20):
        counts, content = _as_host_tensor(counts), _as_host_tensor(content)
        cdt = _TORCH2NP[counts.dtype]
        if not issubclass(cdt.type, numpy.integer):
            raise TypeError("counts must have integer dtype")                      # jagged.py:158-159
        dev = runtime.device()
        self._nevents = int(counts.numel())
        self._nchunks = max(1, min(int(nchunks), max(self._nevents, 1)))
        compute = torch.cuda.current_stream()
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._copy_stream.wait_stream(compute)
        # -- queue every copy up front: counts first, then the content in pieces, one event per piece
        counts_dev = torch.empty(counts.shape, dtype=counts.dtype, device=dev)
        content_dev = torch.empty(content.shape, dtype=content.dtype, device=dev)
        counts_dev.record_stream(self._copy_stream)      # filled on the side stream: keep the allocator informed
        content_dev.record_stream(self._copy_stream)
        per_piece = max(1, int(piece_bytes) // max(1, content.element_size()))
        self._piece_items = per_piece
        self._piece_done = []
        with torch.cuda.stream(self._copy_stream):
            counts_dev.copy_(counts, non_blocking=True)
            counts_landed = torch.cuda.Event()
            counts_landed.record(self._copy_stream)
            for lo in range(0, int(content.numel()), per_piece):
                hi = min(int(content.numel()), lo + per_piece)
                content_dev[lo:hi].copy_(content[lo:hi], non_blocking=True)
                done = torch.cuda.Event()
                done.record(self._copy_stream)
                self._piece_done.append(done)
        self._host_refs = (counts, content)          # the pinned sources must outlive the queued copies
        # -- offsets as soon as the counts are in (the content is still on the bus)
        compute.wait_event(counts_landed)
        self._counts = DeviceArray(counts_dev, cdt)
        self._content = DeviceArray(content_dev, _TORCH2NP[content.dtype])
        offsets, total, status = JaggedArray._scan_counts(self._counts)
        if status & _lib.ST_NEGATIVE_COUNT:
            raise ValueError("counts must be a non-negative array")                # jagged.py:160-161
        if total > len(self._content):
            raise ValueError("maximum offset {0} is beyond the length of the content ({1})".format(total, len(self._content)))
        self._offsets, self._total = offsets, total
        # -- element boundaries of the event chunks: one tiny gather + read-back
        n, k = self._nevents, self._nchunks
        self._event_bounds = [(n * i) // k for i in range(k + 1)]
        where = torch.tensor(self._event_bounds, dtype=torch.int64, device=dev)
        self._element_bounds = [int(x) for x in offsets.tensor[where].cpu().tolist()]

    def __len__(self):
        return self._nevents

    @property
    def nchunks(self):
        return self._nchunks

    def _wait_for_elements(self, upto):
        """Make the compute stream wait until content[:upto] has landed."""
        if upto <= 0 or not self._piece_done:
            return
        piece = min(len(self._piece_done) - 1, (upto - 1) // self._piece_items)
        torch.cuda.current_stream().wait_event(self._piece_done[piece])

    def chunk(self, i):
        """Events [bounds[i], bounds[i+1]) as a view of the resident array (offsets[0] != 0 for i > 0)."""
        lo, hi = self._event_bounds[i], self._event_bounds[i + 1]
        first, last = self._element_bounds[i], self._element_bounds[i + 1]
        self._wait_for_elements(last)
        out = JaggedArray._make(self._offsets[lo:hi + 1], self._content, last, first)
        out._counts = self._counts[lo:hi]
        return out

    def __iter__(self):
        for i in range(self._nchunks):
            yield self.chunk(i)

    def whole(self):
        """The complete array (waits for the last piece)."""
        self._wait_for_elements(self._total)
        out = JaggedArray._make(self._offsets, self._content, self._total, 0)
        out._counts = self._counts
        return out
