"""Device-resident 1-D arrays and the per-process runtime (workspace, stream, status read-back).

PyTorch is plumbing here: it owns device memory (caching allocator), the current CUDA stream and -- in
sharded.py -- the NCCL communicator.  All arithmetic goes through libjagged_sm100a.so (see _lib.py).
"""
import ctypes
import warnings

import numpy
import numpy.lib.mixins

from . import _lib
from ._lib import lib

try:
    import torch
except ImportError as err:          # pragma: no cover
    raise ImportError("awkward0_b200 needs PyTorch for device memory and streams") from err


# ---- dtype tables ---------------------------------------------------------------------------------------------
_NP2JG = {
    numpy.dtype(numpy.float32): _lib.JG_F32, numpy.dtype(numpy.float64): _lib.JG_F64,
    numpy.dtype(numpy.int32): _lib.JG_I32, numpy.dtype(numpy.int64): _lib.JG_I64,
    numpy.dtype(numpy.bool_): _lib.JG_B8, numpy.dtype(numpy.uint8): _lib.JG_U8,
    numpy.dtype(numpy.int8): _lib.JG_I8, numpy.dtype(numpy.int16): _lib.JG_I16,
    numpy.dtype(numpy.uint16): _lib.JG_U16, numpy.dtype(numpy.uint32): _lib.JG_U32,
    numpy.dtype(numpy.uint64): _lib.JG_U64,
}
_NP2TORCH = {
    numpy.dtype(numpy.float32): torch.float32, numpy.dtype(numpy.float64): torch.float64,
    numpy.dtype(numpy.int32): torch.int32, numpy.dtype(numpy.int64): torch.int64,
    numpy.dtype(numpy.bool_): torch.bool, numpy.dtype(numpy.uint8): torch.uint8,
    numpy.dtype(numpy.int8): torch.int8, numpy.dtype(numpy.int16): torch.int16,
    numpy.dtype(numpy.uint16): torch.uint16, numpy.dtype(numpy.uint32): torch.uint32,
    numpy.dtype(numpy.uint64): torch.uint64,
}
_TORCH2NP = dict((v, k) for k, v in _NP2TORCH.items())

INDEXTYPE = numpy.dtype(numpy.int64)
BOOLTYPE = numpy.dtype(numpy.bool_)


def jg_code(dtype):
    dtype = numpy.dtype(dtype)
    try:
        return _NP2JG[dtype]
    except KeyError:
        raise NotImplementedError("dtype {0} is not supported on the device (awkward0_b200 has no CPU fallback)".format(dtype))


def torch_dtype(dtype):
    dtype = numpy.dtype(dtype)
    try:
        return _NP2TORCH[dtype]
    except KeyError:
        raise NotImplementedError("dtype {0} is not supported on the device".format(dtype))


# ---- runtime ----------------------------------------------------------------------------------------------------
class _Runtime(object):
    """Per-(device, stream) scratch: the scan workspace and the small status/total words kernels write."""

    def __init__(self):
        self._ws = {}
        self._scal = {}
        self._side = {}         # per device: (side stream, pinned int64[8]) of read_early
        self._pinned = {}       # per device: pinned int64[8] of read

    _has_cuda = None
    _devices = {}

    @classmethod
    def device(cls):
        if cls._has_cuda is None:
            cls._has_cuda = bool(torch.cuda.is_available())     # checked once: it walks the environment
        if not cls._has_cuda:
            raise RuntimeError("no CUDA device visible: awkward0_b200 runs on B200 only and has no CPU fallback")
        idx = torch.cuda.current_device()
        dev = cls._devices.get(idx)
        if dev is None:
            dev = cls._devices[idx] = torch.device("cuda", idx)
        return dev

    @staticmethod
    def stream():
        """The current CUDA stream of the current device as a cudaStream_t (raw handle, no Stream object)."""
        try:
            return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
        except AttributeError:          # pragma: no cover  (older torch)
            return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _key(self, dev):
        """Scratch is per (device, stream): kernels of two streams may run concurrently, and the descriptors of a scan
        or the zero fill of a fresh scalars pool are only ordered on the stream that issued them."""
        return (dev.index, self.stream().value or 0)

    def workspace(self, n):
        """(pointer, bytes) of a scratch buffer large enough for any call on n events."""
        dev = self.device()
        need = int(lib.jg_workspace_bytes(int(n)))
        key = self._key(dev)
        buf = self._ws.get(key)
        if buf is None or buf.numel() < need:
            buf = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=dev)
            self._ws[key] = buf
        return ctypes.c_void_p(buf.data_ptr()), ctypes.c_size_t(buf.numel())

    def elem_workspace(self, content_len, itemsize):
        """(pointer, bytes) of scratch for the element-tiled calls (sized by content elements, not by events)."""
        dev = self.device()
        need = int(lib.jg_elem_workspace_bytes(int(content_len), int(itemsize)))
        key = self._key(dev) + ("elem",)
        buf = self._ws.get(key)
        if buf is None or buf.numel() < need:
            buf = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=dev)
            self._ws[key] = buf
        return ctypes.c_void_p(buf.data_ptr()), ctypes.c_size_t(buf.numel())

    def scalars(self):
        """A zeroed int64[8] on the device: [0] total, [1] status (low 32 bits), [2:6] info.

        Slices of a pre-zeroed pool (one fill per 512 operations instead of one torch fill kernel per
        operation); a slice is handed out once, so pending read-backs never alias.
        """
        dev = self.device()
        key = self._key(dev)
        pool = self._scal.get(key)
        if pool is None or pool[1] >= pool[0].numel():
            pool = [torch.zeros(8 * 512, dtype=torch.int64, device=dev), 0]
            self._scal[key] = pool
        out = pool[0][pool[1]:pool[1] + 8]
        pool[1] += 8
        return out

    def read_early(self, scal):
        """Start reading `scal` back NOW, on a side stream that waits only for what has been queued so far: returns a
        function that gives (total, status, info) like read() but waits for that copy alone -- kernels queued on the
        compute stream in the meantime keep running while the host goes on (the selection reads K this way after its
        counting pass and queues the compaction before asking for it)."""
        dev = self.device()
        key = dev.index
        side = self._side.get(key)
        if side is None:
            side = self._side[key] = (torch.cuda.Stream(device=dev), torch.empty(8, dtype=torch.int64).pin_memory())
        stream, host = side
        ready = torch.cuda.Event()
        ready.record()                                   # the compute stream, after everything queued so far
        stream.wait_event(ready)
        with torch.cuda.stream(stream):
            host.copy_(scal, non_blocking=True)
            done = torch.cuda.Event()
            done.record(stream)

        def result():
            done.synchronize()
            h = host.numpy()
            return int(h[0]), int(h[1]) & 0xffffffff, h[2:6].copy()
        return result

    def read(self, scal):
        """One device->host read (synchronises the stream): returns (total, status, info[4]).

        Through a PINNED host word per device and an asynchronous copy followed by a stream synchronisation: `.cpu()`
        goes through pageable memory (a staging copy and a blocking call, ~0.13 ms per read-back on the bench host
        against the ~15 us a pinned copy takes; a sweep step makes three of them, profiles/r2_pyprofile_sweep_after.txt).
        """
        dev = self.device()
        host = self._pinned.get(dev.index)
        if host is None:
            host = self._pinned[dev.index] = torch.empty(8, dtype=torch.int64).pin_memory()
        host.copy_(scal, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        h = host.numpy()
        return int(h[0]), int(h[1]) & 0xffffffff, h[2:6].copy()


runtime = _Runtime()


def ptr(t, offset_items=0):
    """Raw device pointer of a torch tensor (optionally advanced by whole items)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr() + offset_items * t.element_size())


def empty(n, dtype):
    return torch.empty(int(n), dtype=torch_dtype(dtype), device=runtime.device())


# ---- DeviceArray --------------------------------------------------------------------------------------------------
class DeviceArray(numpy.lib.mixins.NDArrayOperatorsMixin):
    """A 1-D array resident in HBM with a numpy-like surface (dtype, len, slicing views, tolist, ufuncs).

    It stands where the reference holds a ``numpy.ndarray`` (starts/stops/offsets/content/counts/parents and
    every flat result such as ``a.sum()``).  Arithmetic on it runs the same CUDA kernels as JaggedArray content.
    """

    __array_priority__ = 50
    __slots__ = ("_t", "_dtype")

    def __init__(self, tensor, dtype=None):
        if not isinstance(tensor, torch.Tensor):
            raise TypeError("DeviceArray wraps a torch.Tensor; use asdevice() to convert")
        if tensor.dim() != 1:
            raise ValueError("DeviceArray is one-dimensional")
        if not tensor.is_cuda:
            raise ValueError("DeviceArray must live on a CUDA device (no CPU fallback)")
        if not tensor.is_contiguous():
            tensor = tensor.contiguous()
        self._t = tensor
        self._dtype = numpy.dtype(dtype) if dtype is not None else _TORCH2NP[tensor.dtype]

    # -- numpy-like surface
    @property
    def dtype(self):
        return self._dtype

    @property
    def shape(self):
        return (int(self._t.shape[0]),)

    @property
    def ndim(self):
        return 1

    @property
    def size(self):
        return int(self._t.shape[0])

    @property
    def nbytes(self):
        return int(self._t.numel() * self._t.element_size())

    @property
    def itemsize(self):
        return int(self._t.element_size())

    def __len__(self):
        return int(self._t.shape[0])

    @property
    def tensor(self):
        """The underlying torch tensor (zero-copy)."""
        return self._t

    def data_ptr(self, offset_items=0):
        return ptr(self._t, offset_items)

    def get(self):
        """Copy to a host numpy array (device -> host)."""
        if self._t.dtype in (torch.uint16, torch.uint32, torch.uint64):
            return self._t.view(_NP2TORCH[numpy.dtype("i%d" % self.itemsize)]).cpu().numpy().view(self._dtype)
        return self._t.cpu().numpy().astype(self._dtype, copy=False)

    def __array__(self, dtype=None, copy=None):
        out = self.get()
        return out if dtype is None else out.astype(dtype, copy=False)

    def tolist(self):
        return self.get().tolist()

    def copy(self):
        return DeviceArray(self._t.clone(), self._dtype)

    def view(self, dtype):
        dtype = numpy.dtype(dtype)
        if dtype.itemsize != self._dtype.itemsize:
            raise ValueError("DeviceArray.view needs the same itemsize")
        return DeviceArray(self._t.view(torch_dtype(dtype)), dtype)

    def astype(self, dtype, copy=True):
        dtype = numpy.dtype(dtype)
        if dtype == self._dtype:
            return self.copy() if copy else self
        from . import _ufunc
        return _ufunc.cast(self, dtype)

    def __repr__(self):
        n = len(self)
        if n <= 8:
            body = " ".join(str(x) for x in self.tolist())
        else:
            body = " ".join(str(x) for x in self[:3].tolist()) + " ... " + " ".join(str(x) for x in self[-3:].tolist())
        return "<DeviceArray [{0}] {1} at cuda:{2}>".format(body, self._dtype, self._t.device.index)

    def __getitem__(self, where):
        if isinstance(where, slice):
            start, stop, step = where.indices(len(self))
            if step != 1:
                raise NotImplementedError("DeviceArray slices must have step 1")
            return DeviceArray(self._t[start:max(start, stop)], self._dtype)
        if isinstance(where, (int, numpy.integer)):
            n = len(self)
            i = int(where)
            if i < 0:
                i += n
            if not 0 <= i < n:
                raise IndexError("index {0} is out of bounds for axis 0 with size {1}".format(where, n))
            return self[i:i + 1].get()[0]
        if isinstance(where, (list, numpy.ndarray)):
            where = asdevice(numpy.asarray(where))
        if isinstance(where, DeviceArray) and where.dtype.kind in "iu":
            from . import _ufunc
            return _ufunc.take(self, where)
        if isinstance(where, DeviceArray) and where.dtype == numpy.bool_:
            from . import _ufunc
            return _ufunc.compress([self], where)[0]
        raise NotImplementedError("DeviceArray supports integer, slice, integer-array and boolean-array indexing only")

    def __bool__(self):
        if len(self) == 1:
            return bool(self.get()[0])
        raise ValueError("The truth value of an array with more than one element is ambiguous. Use a.any() or a.all()")

    # -- ufuncs and reductions run on the device
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        from . import _ufunc
        for x in inputs:
            if type(x).__name__ in ("JaggedArray", "Table"):     # the jagged operand drives the broadcast, a Table its columns
                return NotImplemented
        if "out" in kwargs:
            raise NotImplementedError("in-place operations not supported")
        if method != "__call__":
            return NotImplemented
        return _ufunc.flat_ufunc(ufunc, inputs)

    def _reduce(self, op):
        from . import _ufunc
        return _ufunc.flat_reduce(self, op)

    def sum(self):
        return self._reduce(_lib.SUM)

    def device_sum(self):
        """sum() left in HBM as a 1-element DeviceArray (no device-to-host copy, no synchronisation): what a step
        hands to sharded.DeviceTotals so that all its whole-array totals travel in one collective and one read-back."""
        from . import _ufunc
        return _ufunc.flat_reduce_device(self, _lib.SUM)

    def device_max(self):
        from . import _ufunc
        return _ufunc.flat_reduce_device(self, _lib.MAX)

    def device_min(self):
        from . import _ufunc
        return _ufunc.flat_reduce_device(self, _lib.MIN)

    def prod(self):
        return self._reduce(_lib.PROD)

    def min(self):
        return self._reduce(_lib.MIN)

    def max(self):
        return self._reduce(_lib.MAX)

    def any(self):
        return self._reduce(_lib.ANY)

    def all(self):
        return self._reduce(_lib.ALL)


class DeviceMatrix(object):
    """A rectangular (events x count) view of a DeviceArray: what JaggedArray.regular() returns (jagged.py:1053-1068
    returns a 2-D numpy array).  Zero-copy; `.get()` / numpy.asarray() copy to the host as a 2-D array."""

    __slots__ = ("_flat", "_shape")

    def __init__(self, flat, shape):
        if int(shape[0]) * int(shape[1]) != len(flat):
            raise ValueError("cannot reshape array of size {0} into shape {1}".format(len(flat), tuple(shape)))
        self._flat, self._shape = flat, (int(shape[0]), int(shape[1]))

    shape = property(lambda self: self._shape)
    ndim = property(lambda self: 2)
    dtype = property(lambda self: self._flat.dtype)
    flat = property(lambda self: self._flat)

    @property
    def tensor(self):
        return self._flat.tensor.view(self._shape)

    def __len__(self):
        return self._shape[0]

    def get(self):
        return self._flat.get().reshape(self._shape)

    def __array__(self, dtype=None, copy=None):
        out = self.get()
        return out if dtype is None else out.astype(dtype, copy=False)

    def tolist(self):
        return self.get().tolist()

    def __repr__(self):
        return "<DeviceMatrix {0} {1} at cuda>".format(self._shape, self.dtype)


def asdevice(x, dtype=None):
    """Anything array-like -> DeviceArray (host data is copied host->device once)."""
    if isinstance(x, DeviceArray):
        return x if dtype is None or numpy.dtype(dtype) == x.dtype else x.astype(dtype)
    if isinstance(x, torch.Tensor):
        if x.dim() != 1:
            raise ValueError("only one-dimensional device tensors are supported")
        if not x.is_cuda:
            x = x.to(runtime.device())
        out = DeviceArray(x)
        return out if dtype is None or numpy.dtype(dtype) == out.dtype else out.astype(dtype)
    arr = numpy.asarray(x)
    if dtype is not None:
        arr = arr.astype(dtype, copy=False)
    if arr.ndim != 1:
        raise NotImplementedError("only one-dimensional arrays are supported on the device (got shape {0})".format(arr.shape))
    np_dtype = arr.dtype
    tdt = torch_dtype(np_dtype)
    arr = numpy.ascontiguousarray(arr)
    with warnings.catch_warnings():
        # read-only host buffers (Arrow, memory maps) are only ever the SOURCE of the host->device copy
        warnings.simplefilter("ignore", UserWarning)
        if np_dtype.kind == "u" and np_dtype.itemsize > 1:
            host = torch.from_numpy(arr.view("i%d" % np_dtype.itemsize)).view(tdt)
        else:
            host = torch.from_numpy(arr)
    return DeviceArray(host.to(runtime.device()), np_dtype)
