"""GPU-resident JaggedArray with the awkward0 API (reference: awkward0/array/jagged.py, awkward0/array/base.py).

Same method names, argument meaning and exceptions as ``awkward0.JaggedArray`` for the hot path named in
BASELINE.json; every array operation is a hand-written sm_100a kernel reached through the C ABI of
libjagged_sm100a.so (include/jagged_b200.h).  Buffers stay in HBM between calls: starts / stops / offsets /
content / counts / parents and every flat result are ``DeviceArray`` objects (device.py); only ``tolist()``,
``__array__``/``.get()`` and error checks move data to the host.  There is no CPU fallback: operations outside
the hot path (last paragraph) raise NotImplementedError.

Design differences from the reference that do not change results:
  * the canonical layout is dense int64 ``offsets`` (an explicit flag instead of the pointer-alias test of
    jagged.py:25-40); general starts/stops are accepted and compacted on first use (jagged.py:883-942);
  * validity (jagged.py:469-477) is one fused kernel whose flags are read back together with offsets[0],
    offsets[-1]; arrays produced by this library's own kernels are born valid;
  * results whose reference layout is non-dense (argmin/argmax, jagged.py:1597-1600) are returned dense -- the
    logical value (tolist(), counts, flatten()) is identical.
"""
import ctypes
import os
import numbers
from collections import OrderedDict

import numpy
import numpy.lib.mixins

from . import _lib, _ufunc
from . import lazy as _lazy
from ._lib import call
from .device import (BOOLTYPE, INDEXTYPE, DeviceArray, asdevice, empty, jg_code, runtime, torch, torch_dtype)
from .masked import MaskedArray, fill_masked
from .device import DeviceMatrix
from .table import Table

_ELEM_DTYPES = (numpy.dtype(numpy.float32), numpy.dtype(numpy.float64), numpy.dtype(numpy.int32), numpy.dtype(numpy.int64))


def _vp(t, offset_items=0):
    return ctypes.c_void_p(t.data_ptr() + offset_items * t.element_size()) if t is not None else ctypes.c_void_p(0)


def _isintegertype(dtype):
    return issubclass(numpy.dtype(dtype).type, numpy.integer)


class _bothmethod(object):
    """awkward0.util.bothmethod: callable on the class (arrays = all inputs) or on an instance (self comes first)."""

    def __init__(self, fcn):
        self.fcn = fcn

    def __get__(self, ins, typ=None):
        fcn = self.fcn
        if ins is None:
            return lambda *args, **kwargs: fcn(True, typ, *args, **kwargs)
        return lambda *args, **kwargs: fcn(False, ins, *args, **kwargs)


# Selection schedule of the event-tiled path: False = counting pass, scan of the tile totals, compaction (K is read back
# while the compaction runs); True = ONE pass on published tile aggregates (include/jagged_b200.h, phase 3).
SELECT_SINGLE_PASS = [os.environ.get("JG_SELECT_SINGLE", "0") == "1"]
# the fifth single-pass form (ranges of 8 tiles per CTA, include/jagged_b200.h: phase 4), float32 compared columns only
SELECT_RANGES = [os.environ.get("JG_SELECT_RANGES", "0") == "1"]


class JaggedArray(numpy.lib.mixins.NDArrayOperatorsMixin):
    """JaggedArray(starts, stops, content) on the device -- drop-in for awkward0.JaggedArray's hot path."""

    # global switches, same names and defaults as base.py:28-31
    allow_tonumpy = True
    allow_iter = True
    check_prop_valid = True
    check_whole_valid = True
    # argmin/argmax also keep the extreme values (same pass, +v bytes per event) so that a[a.argmax()] is free
    cache_arg_values = True

    DEFAULTTYPE = numpy.dtype(numpy.float64)      # base.py:42-48
    INDEXTYPE = INDEXTYPE
    BOOLTYPE = BOOLTYPE
    MASKTYPE = BOOLTYPE

    __array_priority__ = 100

    # class indirection (base.py:306-374): subclasses can substitute their own node classes
    @property
    def JaggedArray(self):
        return JaggedArray

    @property
    def Table(self):
        return Table

    # ------------------------------------------------------------------------------------------------------
    # classmethods that ARE the reference's kernels
    # ------------------------------------------------------------------------------------------------------
    @classmethod
    def counts2offsets(cls, counts):
        """jagged.py:42-47 -> jg_counts2offsets (single-pass decoupled-lookback scan).  Returns int64[N+1]."""
        offsets, _, _ = cls._scan_counts(asdevice(counts), sync=False)
        return offsets

    @classmethod
    def _scan_counts(cls, counts, sync=True, tiles=False):
        if counts.dtype == numpy.int64 or counts.dtype == numpy.int32:
            src = counts
        elif counts.dtype == numpy.bool_ or counts.dtype == numpy.uint8:
            src = counts
        elif _isintegertype(counts.dtype):
            src = counts.astype(numpy.int64)
        else:
            raise TypeError("counts must have integer dtype")
        n = len(src)
        offsets = empty(n + 1, INDEXTYPE)
        scal = runtime.scalars()
        ws, wsb = runtime.workspace(n)
        if tiles and n > 0 and src.dtype.itemsize >= 4:
            # the same pass also counts the non-empty events of every 512-event tile (what a dense argmin / argmax
            # needs to place its results): jg_segargminmax_dense_tiles then skips its own pass over the offsets
            tile_nonempty = empty((n + 511) // 512, INDEXTYPE)
            call("jg_counts2offsets_tiles", src.data_ptr(), jg_code(src.dtype), n, _vp(offsets), _vp(tile_nonempty),
                 _vp(scal), _vp(scal, 1), ws, wsb, runtime.stream())
        else:
            tile_nonempty = None
            call("jg_counts2offsets", src.data_ptr(), jg_code(src.dtype), n, _vp(offsets), _vp(scal), _vp(scal, 1), ws,
                 wsb, runtime.stream())
        total, status = None, 0
        if sync:
            total, status, _ = runtime.read(scal)
        if tiles:
            return DeviceArray(offsets, INDEXTYPE), total, status, tile_nonempty
        return DeviceArray(offsets, INDEXTYPE), total, status

    @classmethod
    def offsets2parents(cls, offsets):
        """jagged.py:49-54 -> jg_offsets2parents."""
        offsets = cls._as_index64(asdevice(offsets))
        n = len(offsets) - 1
        if n < 0:
            raise ValueError("offsets must be a non-empty array")
        first, last = (int(x) for x in offsets.tensor[[0, -1]].cpu().tolist())
        return cls._parents_from(offsets, n, last, first)

    # Count distribution -> kernel variant (north_star: "variants chosen by count distribution").  The event tile (512
    # events, 16 KB of staged content) fits mean counts up to ~7; beyond that -- Poisson(40), a few huge events -- the
    # element-tiled kernels of csrc/jg_elem.cu split the work by CONTENT instead.  In between the event-tiled kernels are
    # kept where a larger tile pays (measured per operation: _stage_hint for the reductions, a factor of 2 for parents /
    # localindex / broadcast, whose owner table covers 8192 positions).  Every variant gives the same results; a tile
    # that does not fit after all (one long event among short ones) is handled on the device, per tile.
    LONG_EVENTS_MEAN = 7

    @classmethod
    def _long_events(cls, n, content_len, wide=1):
        return n > 0 and content_len > cls.LONG_EVENTS_MEAN * wide * n and content_len >= (1 << 14)

    @classmethod
    def _stage_hint(cls, n, content_len, dtype, arg=False):
        """(element tiles?, hint for the event-tiled call).  The reductions keep 4-byte contents with mean counts of
        7 - 14 on the event-tiled kernels by staging 32 KB per tile (JG_WIDE_STAGE) and, where the tiles are full enough
        to pay for 3 CTAs per SM, 64 KB (JG_WIDER_STAGE: means of 14 - 28 for the arg-reductions, 20 - 28 for the others;
        measured on Poisson(15 / 20 / 26), profiles/r2_exp43.log); elsewhere -- and from 7 on for 8-byte items, whose
        stage is 32 KB already -- the element-tiled kernels."""
        dtype = numpy.dtype(dtype)
        can_elem = dtype in _ELEM_DTYPES
        if dtype.itemsize != 4:
            return (can_elem and cls._long_events(n, content_len)), 0
        if not cls._long_events(n, content_len):                    # mean <= 7
            return False, 0
        if not cls._long_events(n, content_len, 2):                 # 7 < mean <= 14
            return False, _lib.WIDE_STAGE
        if cls._long_events(n, content_len, 4):                     # mean > 28
            return can_elem, (0 if can_elem else _lib.WIDER_STAGE)
        if arg or not can_elem or cls._long_events(n, content_len, 20.0 / 7.0):
            return False, _lib.WIDER_STAGE
        return True, 0

    @classmethod
    def _parents_from(cls, off64, n, last, first=None):
        out = empty(last, INDEXTYPE)
        if last > 0:
            if first is not None and cls._long_events(n, last - first, 2):      # (the event-tiled table covers means up to 14)
                ws, wsb = runtime.elem_workspace(last - first, 4)
                call("jg_expand_elem", 0, ctypes.c_void_p(0), 8, off64.data_ptr(), n, _vp(out), last - first, ws, wsb,
                     runtime.stream())
            else:
                call("jg_offsets2parents", off64.data_ptr(), n, _vp(out), runtime.stream())
        return DeviceArray(out, INDEXTYPE)

    @classmethod
    def startsstops2parents(cls, starts, stops):
        """jagged.py:56-71 -> jg_startsstops2parents."""
        starts = cls._as_index64(asdevice(starts))
        stops = cls._as_index64(asdevice(stops))
        n = len(starts)
        out_len = int(stops[:n].max()) if n > 0 else 0
        out = empty(out_len, INDEXTYPE)
        call("jg_startsstops2parents", starts.data_ptr(), stops.data_ptr(), n, _vp(out), out_len, runtime.stream())
        return DeviceArray(out, INDEXTYPE)

    @staticmethod
    def _as_index64(x):
        if x.dtype == numpy.int64:
            return x
        if not _isintegertype(x.dtype):
            raise TypeError("offsets must have integer dtype")
        return x.astype(numpy.int64)

    @classmethod
    def offsetsaliased(cls, starts, stops):
        """jagged.py:25-40 -- are starts/stops the [:-1] / [1:] views of one offsets buffer?"""
        if isinstance(starts, DeviceArray) and isinstance(stops, DeviceArray):
            a, b = starts.tensor, stops.tensor
            return (a.dtype == b.dtype and len(starts) == len(stops) and
                    a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr() and
                    b.data_ptr() == a.data_ptr() + a.element_size() and
                    a.untyped_storage().nbytes() >= (a.storage_offset() + len(starts) + 1) * a.element_size())
        if isinstance(starts, numpy.ndarray) and isinstance(stops, numpy.ndarray):
            return (starts.base is not None and starts.base is stops.base and starts.ndim == 1 and stops.ndim == 1 and
                    isinstance(starts.base, numpy.ndarray) and starts.base.ndim == 1 and
                    starts.dtype == stops.dtype == starts.base.dtype and
                    starts.ctypes.data == starts.base.ctypes.data and
                    stops.ctypes.data == starts.base.ctypes.data + stops.dtype.itemsize and
                    len(starts) == len(starts.base) - 1 and len(stops) == len(stops.base) - 1)
        return False

    # ------------------------------------------------------------------------------------------------------
    # construction
    # ------------------------------------------------------------------------------------------------------
    def __init__(self, starts, stops, content):
        self._reset()
        self._content = self._tocontent(content)
        if self.offsetsaliased(starts, stops):
            if isinstance(starts, DeviceArray):
                t = starts.tensor
                base = DeviceArray(torch.as_strided(t, (len(starts) + 1,), (1,), t.storage_offset()), starts.dtype)
            else:
                base = asdevice(starts.base)
            self._set_offsets(base)
        else:
            starts = asdevice(starts if not isinstance(starts, (list, tuple)) or len(starts) else numpy.zeros(0, INDEXTYPE))
            stops = asdevice(stops if not isinstance(stops, (list, tuple)) or len(stops) else numpy.zeros(0, INDEXTYPE))
            if self.check_prop_valid:
                if not _isintegertype(starts.dtype):
                    raise TypeError("starts must have integer dtype")
                if not _isintegertype(stops.dtype):
                    raise TypeError("stops must have integer dtype")
            if len(starts) > len(stops):
                raise ValueError("starts must have the same (or shorter) length than stops")
            self._starts, self._stops = starts, stops[:len(starts)]
            self._check_general()

    def _reset(self):
        self._starts = self._stops = self._offsets = self._off64 = None
        self._counts = self._parents = None
        self._isvalid = False
        self._first = self._last = None      # host copies of offsets[0], offsets[-1]
        self._pending_status = 0
        self._rebased = None
        self._general = None
        self._maxcount = None
        self._tile_nonempty = None       # non-empty events per 512-event tile, a by-product of fromcounts' scan
        self._arg_source = None          # set on argmin/argmax results: (offsets, content, write epoch) they were computed from
        self._arg_values = None
        self._nestedcross = None         # set on cross(nested=True) results: the flat form (jagged.py:1357)

    @classmethod
    def _tocontent(cls, content):
        if isinstance(content, (DeviceArray, Table, JaggedArray, MaskedArray)):
            return content
        if isinstance(content, torch.Tensor):
            return asdevice(content)
        arr = numpy.asarray(content)
        if arr.dtype == numpy.dtype(object) or arr.size == 0 and not isinstance(content, numpy.ndarray):
            arr = numpy.asarray(content, dtype=cls.DEFAULTTYPE)
        return asdevice(arr)

    def _set_offsets(self, offsets):
        """Adopt a dense offsets buffer (user dtype kept for the properties, int64 copy for the kernels)."""
        if not _isintegertype(offsets.dtype):
            raise TypeError("offsets must have integer dtype")
        if len(offsets) == 0:
            raise ValueError("offsets must be a non-empty array")
        self._offsets = offsets
        self._off64 = self._as_index64(offsets)
        self._starts, self._stops = offsets[:-1], offsets[1:]
        self._counts = self._parents = self._rebased = None
        self._isvalid = False
        self._first = self._last = None
        if self.check_prop_valid or self.check_whole_valid:
            self._run_validate()
            if self.check_prop_valid and self._pending_status & _lib.ST_NEGATIVE_OFFSET:
                raise ValueError("offsets must be a non-negative array")

    def _content_len(self):
        return len(self._content)

    def _run_validate(self):
        n = len(self._off64) - 1
        scal = runtime.scalars()
        call("jg_validate_offsets", self._off64.data_ptr(), n, self._content_len(), _vp(scal, 2), _vp(scal, 1),
             runtime.stream())
        _, status, info = runtime.read(scal)
        self._pending_status = status
        self._first, self._last = int(info[0]), int(info[1])
        self._maxcount = int(info[2])

    def _check_general(self):
        """General starts/stops: validate, and adopt dense offsets when starts[1:] == stops[:-1] (jagged.py:358)."""
        n = len(self._starts)
        s64, e64 = self._as_index64(self._starts), self._as_index64(self._stops)
        scal = runtime.scalars()
        call("jg_validate_startsstops", s64.data_ptr(), e64.data_ptr(), n, self._content_len(), _vp(scal, 2),
             _vp(scal, 1), runtime.stream())
        _, status, info = runtime.read(scal)
        if self.check_prop_valid and status & _lib.ST_NEGATIVE_OFFSET:
            raise ValueError("starts must be a non-negative array")
        self._pending_status = status
        self._general = (s64, e64)
        if not status & _lib.ST_NOT_OFFSETS and not status & _lib.ST_STOPS_LT_STARTS:
            # offsets-compatible: build the dense offsets once (append(starts, stops[-1]), jagged.py:362)
            off = empty(n + 1, INDEXTYPE)
            if n > 0:
                off[:n].copy_(s64.tensor)
                off[n:].copy_(e64.tensor[n - 1:])
            else:
                off.zero_()
            self._off64 = DeviceArray(off, INDEXTYPE)
            self._offsets = self._off64 if self._starts.dtype == numpy.int64 else None
            if n > 0:
                first, last = self._off64[0], self._off64[-1]
            else:
                first = last = 0
            self._first, self._last = int(first), int(last)

    @classmethod
    def _make(cls, off64, content, last, first=0):
        """Internal constructor for arrays produced by this library's kernels (dense, valid, bounds known)."""
        out = cls.__new__(cls)
        out._reset()
        out._content = content
        out._offsets = out._off64 = off64
        out._starts, out._stops = off64[:-1], off64[1:]
        out._isvalid = True
        out._first, out._last = first, last
        return out

    @classmethod
    def _from_general(cls, s64, e64, content):
        """Internal constructor: a valid general (starts, stops) layout selected from a valid array."""
        out = cls.__new__(cls)
        out._reset()
        out._content = content
        out._starts, out._stops = s64, e64
        out._general = (s64, e64)
        out._isvalid = True
        return out

    @classmethod
    def fromiter(cls, iterable):
        """Host-side builder (generate.py is out of scope): nested python lists -> counts + content -> device."""
        rows = [list(x) for x in iterable]
        counts = numpy.array([len(x) for x in rows], dtype=INDEXTYPE)
        flat = [y for x in rows for y in x]
        if any(isinstance(y, (list, tuple, numpy.ndarray)) for y in flat):
            return cls.fromcounts(counts, cls.fromiter(flat))           # lists of lists: one more jagged level
        content = numpy.array(flat) if len(flat) else numpy.array([], dtype=cls.DEFAULTTYPE)
        return cls.fromcounts(counts, content)

    @classmethod
    def fromoffsets(cls, offsets, content):
        """jagged.py:142-153."""
        if isinstance(offsets, (list, tuple)):
            offsets = numpy.asarray(offsets) if len(offsets) else numpy.zeros(0, dtype=INDEXTYPE)
        offsets = asdevice(offsets)
        if not _isintegertype(offsets.dtype):
            raise TypeError("offsets must have integer dtype")
        if len(offsets) == 0:
            raise ValueError("offsets must be a non-empty array")
        out = cls.__new__(cls)
        out._reset()
        out._content = cls._tocontent(content)
        out._set_offsets(offsets)
        return out

    @classmethod
    def fromcounts(cls, counts, content):
        """jagged.py:155-166 -> jg_counts2offsets; negative counts / non-integer dtype raise like the reference."""
        if isinstance(counts, (list, tuple)):
            counts = numpy.asarray(counts) if len(counts) else numpy.zeros(0, dtype=INDEXTYPE)
        counts = asdevice(counts)
        if not _isintegertype(counts.dtype) and counts.dtype != numpy.bool_:
            raise TypeError("counts must have integer dtype")
        if counts.dtype == numpy.bool_:
            raise TypeError("counts must have integer dtype")
        content = cls._tocontent(content)
        offsets, total, status, tile_nonempty = cls._scan_counts(counts, tiles=True)
        if status & _lib.ST_NEGATIVE_COUNT:
            raise ValueError("counts must be a non-negative array")
        out = cls._make(offsets, content, total)
        out._tile_nonempty = tile_nonempty
        out._counts = counts
        out._isvalid = False
        out._pending_status = _lib.ST_BEYOND_CONTENT if total > len(content) else 0
        return out

    # ------------------------------------------------------------------------------------------------------
    # the other constructors (jagged.py:73-110, 168-242): compositions of compare / compaction / scatter kernels
    # ------------------------------------------------------------------------------------------------------
    @classmethod
    def parents2startsstops(cls, parents, length=None):
        """jagged.py:73-93: children contiguous, not necessarily in order or covering; negative parents are skipped."""
        parents = cls._as_index64(asdevice(parents))
        n = len(parents)
        if n == 0:
            zero = DeviceArray(empty(0, INDEXTYPE), INDEXTYPE)
            return zero, zero
        edge = _ufunc.binary(numpy.not_equal, parents[1:], parents[:-1])
        inner = _ufunc.binary(numpy.add, _ufunc.nonzero(edge), numpy.int64(1))
        changes = empty(len(inner) + 2, INDEXTYPE)
        changes[0:1].zero_()
        changes[1:len(inner) + 1].copy_(inner.tensor)
        changes[len(inner) + 1:].fill_(n)
        changes = DeviceArray(changes, INDEXTYPE)
        nevents = int(parents.max()) + 1
        starts = DeviceArray(empty(max(nevents, 0), INDEXTYPE).zero_(), INDEXTYPE)
        counts = DeviceArray(empty(max(nevents, 0), INDEXTYPE).zero_(), INDEXTYPE)
        where = _ufunc.take(parents, changes[:-1])
        _ufunc.put(starts, where, changes[:-1], skip_negative=True)
        _ufunc.put(counts, where, _ufunc.binary(numpy.subtract, changes[1:], changes[:-1]), skip_negative=True)
        return starts, _ufunc.binary(numpy.add, starts, counts)

    @classmethod
    def uniques2offsetsparents(cls, uniques):
        """jagged.py:95-110: children contiguous, in order and covering (no empty events); values only mark runs."""
        uniques = asdevice(uniques)
        n = len(uniques)
        if n == 0:
            offsets = asdevice(numpy.zeros(2, dtype=INDEXTYPE))
            return offsets, DeviceArray(empty(0, INDEXTYPE), INDEXTYPE)
        edge = _ufunc.binary(numpy.not_equal, uniques[1:], uniques[:-1])
        inner = _ufunc.binary(numpy.add, _ufunc.nonzero(edge), numpy.int64(1))
        offsets = empty(len(inner) + 2, INDEXTYPE)
        offsets[0:1].zero_()
        offsets[1:len(inner) + 1].copy_(inner.tensor)
        offsets[len(inner) + 1:].fill_(n)
        offsets = DeviceArray(offsets, INDEXTYPE)
        return offsets, cls.offsets2parents(offsets)

    @classmethod
    def fromparents(cls, parents, content, length=None):
        """jagged.py:168-179."""
        parents = asdevice(parents if not isinstance(parents, (list, tuple)) else numpy.asarray(parents, dtype=INDEXTYPE))
        if not _isintegertype(parents.dtype):
            raise TypeError("parents must have integer dtype")
        content = cls._tocontent(content)
        if len(parents) != len(content):
            raise ValueError("parents array must be one-dimensional with the same length as content")
        starts, stops = cls.parents2startsstops(parents, length=length)
        return cls(starts, stops, content)

    @classmethod
    def fromuniques(cls, uniques, content):
        """jagged.py:181-192."""
        uniques = asdevice(uniques if not isinstance(uniques, (list, tuple)) else numpy.asarray(uniques, dtype=INDEXTYPE))
        if not _isintegertype(uniques.dtype):
            raise TypeError("uniques must have integer dtype")
        content = cls._tocontent(content)
        if len(uniques) != len(content):
            raise ValueError("uniques array must be one-dimensional with the same length as content")
        offsets, parents = cls.uniques2offsetsparents(uniques)
        out = cls.fromoffsets(offsets, content)
        out._parents = parents
        return out

    @classmethod
    def fromlocalindex(cls, index, content, validate=True):
        """jagged.py:194-222."""
        original_counts = None
        if isinstance(index, JaggedArray):
            if validate:
                original_counts = index.counts
            index = index.flatten()
        index = asdevice(index if not isinstance(index, (list, tuple)) else numpy.asarray(index, dtype=INDEXTYPE))
        if not _isintegertype(index.dtype):
            raise TypeError("localindex must have integer dtype")
        content = cls._tocontent(content)
        if len(index) != len(content):
            raise ValueError("localindex array must be one-dimensional with the same length as content")
        if validate and len(index) > 1:
            step_ok = _ufunc.binary(numpy.equal, _ufunc.binary(numpy.subtract, index[1:], index[:-1]), 1)
            is_zero = _ufunc.binary(numpy.equal, index[1:], 0)
            if not bool(_ufunc.binary(numpy.logical_or, step_ok, is_zero).all()):
                raise ValueError("every localindex that is not zero must be one greater than the previous")
        starts = _ufunc.nonzero(_ufunc.binary(numpy.equal, index, 0))
        offsets = empty(len(starts) + 1, INDEXTYPE)
        offsets[:len(starts)].copy_(starts.tensor)
        offsets[len(starts):].fill_(len(index))
        offsets = DeviceArray(offsets, INDEXTYPE)
        if original_counts is not None:
            derived = _ufunc.binary(numpy.subtract, offsets[1:], offsets[:-1])
            same = len(derived) == len(original_counts) and bool(_ufunc.binary(numpy.equal, derived, original_counts.astype(INDEXTYPE)).all())
            if not same:
                raise ValueError("jagged structure of index does not match jagged structure derived from localindex")
        return cls.fromoffsets(offsets, content)

    @classmethod
    def fromjagged(cls, jagged):
        """jagged.py:224-226."""
        return cls(jagged.starts, jagged.stops, jagged.content)

    @classmethod
    def fromfolding(cls, content, size):
        """jagged.py:237-242: consecutive events of `size` items, the last one possibly shorter."""
        content = cls._tocontent(content)
        size = int(size)
        quotient = -(-len(content) // size)
        offsets = _ufunc.binary(numpy.multiply, _ufunc.arange(quotient + 1), numpy.int64(size))
        if len(offsets) > 0:
            offsets = _ufunc.binary(numpy.minimum, offsets, numpy.int64(len(content)))
        return cls.fromoffsets(offsets, content)

    @classmethod
    def fromregular(cls, regular):
        """jagged.py:228-235: a rectangular host array becomes nested events of equal length."""
        regular = numpy.asarray(regular)
        if regular.ndim <= 1:
            raise ValueError("regular array must have more than one dimension")
        out = asdevice(numpy.ascontiguousarray(regular).reshape(-1))
        for x in regular.shape[:0:-1]:
            out = cls.fromfolding(out, x)
        return out

    def copy(self, starts=None, stops=None, content=None):
        out = self.__class__.__new__(self.__class__)
        out.__dict__.update(self.__dict__)
        if content is not None or starts is not None or stops is not None:
            out._arg_source = out._arg_values = None
        if content is not None:
            out._content = self._tocontent(content)
            out._isvalid = False
        if starts is not None or stops is not None:
            s = self._starts if starts is None else starts
            e = self._stops if stops is None else stops
            fresh = JaggedArray(s, e, out._content)
            out.__dict__.update(fresh.__dict__)
        return out

    def deepcopy(self, starts=None, stops=None, content=None):
        """jagged.py:261-269: a copy that shares no buffer with this array (same layout)."""
        out = self.copy(starts=starts, stops=stops, content=content)
        out._valid()

        def clone(c):
            if isinstance(c, JaggedArray):
                return c.deepcopy()
            if isinstance(c, Table):
                return c.map(clone)
            return c.copy()

        if out._off64 is not None and out._general is None:
            return self.fromoffsets(out.offsets.copy(), clone(out._content))
        return JaggedArray(out._starts.copy(), out._stops.copy(), clone(out._content))

    def _like(self, make):
        if isinstance(self._content, JaggedArray):
            return self.copy(content=self._content._like(make))
        return self.copy(content=self._map_content(lambda c: DeviceArray(make(c.tensor), c.dtype)))

    def zeros_like(self, **overrides):
        """jagged.py:277-281 (a fill, not arithmetic: NaN * 0 would not do)."""
        return self._like(torch.zeros_like)

    def ones_like(self, **overrides):
        """jagged.py:283-287."""
        return self._like(torch.ones_like)

    def empty_like(self, **overrides):
        """jagged.py:271-275."""
        return self._like(torch.empty_like)

    def astype(self, dtype):
        """base.py: the content converted, the structure shared."""
        if isinstance(self._content, JaggedArray):
            return self.copy(content=self._content.astype(dtype))
        return self.copy(content=self._map_content(lambda c: c.astype(dtype)))

    def structure1d(self, levellimit=None):
        """jagged.py:1431-1454: starts / stops are one-dimensional here by construction."""
        return self.copy()

    def flattentuple(self):
        """jagged.py:1400-1401."""
        if isinstance(self._content, Table):
            return self.copy(content=self._content.flattentuple())
        if isinstance(self._content, JaggedArray):
            return self.copy(content=self._content.flattentuple())
        return self.copy()

    @property
    def size(self):
        return len(self)

    # ------------------------------------------------------------------------------------------------------
    # validity
    # ------------------------------------------------------------------------------------------------------
    def _valid(self):
        """jagged.py:469-492."""
        if self._isvalid:
            return
        if self.check_whole_valid:
            st = self._pending_status
            if self._off64 is not None and self._general is None:
                if self._first is None:
                    self._run_validate()
                    st = self._pending_status
                if st & _lib.ST_NEGATIVE_OFFSET:
                    raise ValueError("offsets must be a non-negative array")
                if st & _lib.ST_NONMONOTONE:
                    raise ValueError("offsets must be monatonically increasing")
                if st & _lib.ST_BEYOND_CONTENT:
                    raise ValueError("maximum offset {0} is beyond the length of the content ({1})".format(self._last, self._content_len()))
            else:
                if st & _lib.ST_STOPS_LT_STARTS:
                    raise ValueError("stops must be greater than or equal to starts")
                if st & _lib.ST_BEYOND_CONTENT:
                    raise ValueError("maximum stop is beyond the length of the content ({0})".format(self._content_len()))
        elif self._off64 is not None and self._first is None:
            n = len(self._off64) - 1
            self._first, self._last = int(self._off64[0]), int(self._off64[n])
        self._isvalid = True

    def valid(self, exception=False, message=False):
        """base.py:174-188."""
        try:
            self._valid()
        except Exception as err:
            if exception:
                raise err
            elif message:
                return "{0}: {1}".format(type(err), str(err))
            else:
                return False
        else:
            return None if message else True

    def _canuseoffset(self):
        self._valid()
        return self._off64 is not None

    @property
    def iscompact(self):
        return len(self._starts) == 0 or self._off64 is not None

    def compact(self):
        """jagged.py:1386-1398 -> counts, scan, jg_compact_gather."""
        self._valid()
        if self._off64 is not None:
            return self
        s64, e64 = self._general
        n = len(s64)
        counts = empty(n, INDEXTYPE)
        call("jg_startsstops2counts", s64.data_ptr(), e64.data_ptr(), n, _vp(counts), runtime.stream())
        offsets, total, _ = self._scan_counts(DeviceArray(counts, INDEXTYPE))
        if isinstance(self._content, JaggedArray):
            # jagged-of-jagged: the inner EVENTS are picked (the list of reachable inner event numbers is the
            # compaction of an arange); the inner array is left in general layout and compacts itself on first use
            picks = self._compact_column(_ufunc.arange(len(self._content)), s64, offsets, n, total)
            content = self._content[picks]
        else:
            content = self._map_content(lambda c: self._compact_column(c, s64, offsets, n, total))
        out = self._make(offsets, content, total)
        out._counts = DeviceArray(counts, INDEXTYPE)
        return out

    @staticmethod
    def _compact_column(col, s64, offsets, n, total):
        out = empty(total, col.dtype)
        if total > 0:
            call("jg_compact_gather", col.data_ptr(), col.itemsize, s64.data_ptr(), offsets.data_ptr(), n, _vp(out),
                 runtime.stream())
        return DeviceArray(out, col.dtype)

    def _map_content(self, fcn, content=None):
        content = self._content if content is None else content
        if isinstance(content, Table):
            return content.map(fcn)
        if isinstance(content, JaggedArray):
            raise NotImplementedError("jagged-of-jagged content is outside the GPU hot path (no CPU fallback)")
        if isinstance(content, MaskedArray):
            raise NotImplementedError("masked content (the result of pad) only supports fillna / regular / tolist on the device; call fillna() first")
        return fcn(content)

    def _dense(self, masked_ok=False):
        """(compact array, off64, nevents, first, last): the form every kernel consumes."""
        if not masked_ok and _has_masked(self._content):
            raise NotImplementedError("masked content (the result of pad) only supports fillna / regular / tolist / counts on the device; call fillna() first")
        a = self.compact()
        a._valid()
        return a, a._off64, len(a._off64) - 1, a._first, a._last

    # ------------------------------------------------------------------------------------------------------
    # properties
    # ------------------------------------------------------------------------------------------------------
    @property
    def starts(self):
        return self._starts

    @property
    def stops(self):
        return self._stops

    @property
    def content(self):
        return self._content

    @property
    def offsets(self):
        """jagged.py:352-365."""
        self._valid()
        if self._off64 is None:
            raise ValueError("starts and stops are not compatible with a single offsets array")
        if self._offsets is None:
            self._offsets = self._off64.astype(self._starts.dtype) if self._starts.dtype != numpy.int64 else self._off64
        return self._offsets

    @property
    def counts(self):
        """jagged.py:383-388 -> jg_offsets2counts."""
        if self._counts is None:
            self._valid()
            n = len(self._starts)
            out = empty(n, INDEXTYPE)
            if self._off64 is not None:
                call("jg_offsets2counts", self._off64.data_ptr(), n, _vp(out), runtime.stream())
            else:
                s64, e64 = self._general
                call("jg_startsstops2counts", s64.data_ptr(), e64.data_ptr(), n, _vp(out), runtime.stream())
            counts = DeviceArray(out, INDEXTYPE)
            if self._starts.dtype != numpy.int64:
                counts = counts.astype(self._starts.dtype)
            self._counts = counts
        return self._counts

    @property
    def parents(self):
        """jagged.py:408-416."""
        if self._parents is None:
            self._valid()
            if self._off64 is not None:
                self._parents = self._parents_from(self._off64, len(self._off64) - 1, self._last, self._first)
            else:
                self._parents = self.startsstops2parents(*self._general)
        return self._parents

    @property
    def localindex(self):
        """jagged.py:430-443 -> jg_localindex."""
        a, off, n, first, last = self._dense()
        out = empty(last - first, INDEXTYPE)
        if last > first:
            if self._long_events(n, last - first, 2):
                ws, wsb = runtime.elem_workspace(last - first, 4)
                call("jg_expand_elem", 1, ctypes.c_void_p(0), 8, off.data_ptr(), n, _vp(out), last - first, ws, wsb,
                     runtime.stream())
            else:
                call("jg_localindex", off.data_ptr(), n, _vp(out), runtime.stream())
        return self._make(a._rebased_offsets(), DeviceArray(out, INDEXTYPE), last - first)

    def _rebased_offsets(self):
        """offsets - offsets[0] (shared, not re-scanned, when offsets[0] == 0)."""
        self._valid()
        if self._first == 0:
            return self._off64
        if self._rebased is None:
            self._rebased = _ufunc.binary(numpy.subtract, self._off64, numpy.int64(self._first))
        return self._rebased

    def _nested_parts(self):
        """(compact self, rebased outer offsets, dense inner JaggedArray, number of inner events)."""
        a, off, n, first, last = self._dense()
        return a, a._rebased_offsets(), a._dense_content(), last - first

    def _dense_content(self, content=None):
        content = self._content if content is None else content
        if self._first == 0 and self._last == len(content):
            return content
        return content[self._first:self._last]

    def __len__(self):
        return len(self._starts)

    @property
    def shape(self):
        return (len(self._starts),)

    @property
    def ndim(self):
        return 1

    @property
    def dtype(self):
        return numpy.dtype(object)

    @property
    def type(self):
        inner = self._content.type if isinstance(self._content, JaggedArray) else getattr(self._content, "dtype", "Table")
        return "[0, {0}) -> [0, inf) -> {1}".format(len(self), inner)

    @property
    def nbytes(self):
        head = self._off64.nbytes if self._off64 is not None else self._starts.nbytes + self._stops.nbytes
        return head + self._content.nbytes

    @property
    def columns(self):
        return self._content.columns if isinstance(self._content, (Table, JaggedArray)) else []

    def __getattr__(self, where):
        if where.startswith("_"):
            raise AttributeError(where)
        content = self.__dict__.get("_content")
        if isinstance(content, Table) and where in content:
            return self[where]
        if isinstance(content, JaggedArray) and where in content.columns:
            return self[where]
        raise AttributeError("no column named {0}".format(repr(where)))

    @property
    def i0(self):
        return self["0"]

    @property
    def i1(self):
        return self["1"]

    @property
    def i2(self):
        return self["2"]

    @property
    def i3(self):
        return self["3"]

    @property
    def i4(self):
        return self["4"]

    @property
    def i5(self):
        return self["5"]

    @property
    def i6(self):
        return self["6"]

    @property
    def i7(self):
        return self["7"]

    @property
    def i8(self):
        return self["8"]

    @property
    def i9(self):
        return self["9"]

    @property
    def rowname(self):
        """base.py:655-682: the row name of the records inside; TypeError when there are none."""
        content = self._content
        if isinstance(content, (Table, JaggedArray)):
            return content.rowname
        raise TypeError("not a Table, so there is no rowname")

    @property
    def istuple(self):
        """base.py:684-687."""
        columns = self.columns
        return self.rowname == "tuple" and columns == [str(x) for x in range(len(columns))]

    def boolmask(self, maskedwhen=True):
        """jagged.py:1902-1906: a JaggedArray masks no event."""
        n = len(self)
        return asdevice(numpy.zeros(n, dtype=self.MASKTYPE) if maskedwhen else numpy.ones(n, dtype=self.MASKTYPE))

    @property
    def ismasked(self):
        return self.boolmask(maskedwhen=True)

    @property
    def isunmasked(self):
        return self.boolmask(maskedwhen=False)

    def unzip(self):
        return tuple(self[n] for n in self.columns)

    @classmethod
    def zip(cls, columns1=None, *columns2, **columns3):
        """Subset of jagged.py:1736-1849: jagged columns with identical counts -> one JaggedArray of records."""
        table = OrderedDict()
        if isinstance(columns1, dict):
            table.update(columns1)
            rowname = "Row"
        elif columns1 is not None:
            rowname = "tuple"
            table["0"] = columns1
            for i, x in enumerate(columns2):
                table[str(i + 1)] = x
        else:
            rowname = "Row"
        table.update(columns3)
        first = None
        for x in table.values():
            if isinstance(x, JaggedArray):
                first = x
                break
        if first is None:
            raise TypeError("zip needs at least one JaggedArray")
        a, off, n, _, last = first._dense()
        contents = Table.named(rowname)
        for name, x in table.items():
            if isinstance(x, JaggedArray):
                x = a._align(x)
                contents[name] = x._dense_content()
            else:
                contents[name] = a.tojagged(x)._dense_content()
        return cls._make(a._rebased_offsets(), contents, last - a._first)

    # ------------------------------------------------------------------------------------------------------
    # host conversions
    # ------------------------------------------------------------------------------------------------------
    def _checktonumpy(self):
        if not self.allow_tonumpy:
            raise RuntimeError("awkward0.array.base.AwkwardArray.allow_tonumpy is False; refusing to convert to Numpy")

    def _host(self):
        self._checktonumpy()
        self._valid()
        if self._off64 is not None:
            off = self._off64.get()
            starts, stops = off[:-1], off[1:]
        else:
            starts, stops = self._general[0].get(), self._general[1].get()
        content = self._content
        if isinstance(content, Table):
            content = content._host_columns()
        elif isinstance(content, JaggedArray):
            content = content.tolist()
        else:
            content = content.get()
        return starts, stops, content

    def tolist(self):
        """base.py:159-172."""
        starts, stops, content = self._host()
        def rows(cols, a, b):
            flat = OrderedDict()
            for n, c in cols.items():
                flat[n] = rows(c, a, b) if isinstance(c, dict) else c[a:b].tolist()
            names = list(flat)
            count = b - a
            tup = getattr(cols, "istuple", False)          # each Table prints its own way (tuple rows / record rows)
            return [tuple(flat[n][i] for n in names) if tup else dict((n, flat[n][i]) for n in names) for i in range(count)]

        out = []
        for a, b in zip(starts.tolist(), stops.tolist()):
            if isinstance(content, dict):
                out.append(rows(content, a, b))
            elif isinstance(content, list):
                out.append(content[a:b])
            else:
                out.append(content[a:b].tolist())
        return out

    def __iter__(self):
        if not self.allow_iter:
            raise RuntimeError("awkward0.array.base.AwkwardArray.allow_iter is False; refusing to iterate")
        starts, stops, content = self._host()
        if isinstance(content, dict):
            raise NotImplementedError("iteration over Table content is not supported on the device")
        for a, b in zip(starts.tolist(), stops.tolist()):
            yield content[a:b]

    def __array__(self, dtype=None, copy=None):
        self._checktonumpy()
        out = numpy.empty(len(self), dtype=object)
        for i, x in enumerate(self):
            out[i] = x
        return out

    def __str__(self):
        n = len(self)
        if n <= 6:
            return "[{0}]".format(" ".join(str(x) for x in self.tolist()))
        head, tail = self[:3].tolist(), self[n - 3:].tolist()
        return "[{0} ... {1}]".format(" ".join(str(x) for x in head), " ".join(str(x) for x in tail))

    def __repr__(self):
        return "<JaggedArray {0} at 0x{1:012x}>".format(str(self), id(self))

    def __bool__(self):
        raise ValueError("The truth value of an array with more than one element is ambiguous. Use a.any() or a.all()")

    # ------------------------------------------------------------------------------------------------------
    # __getitem__
    # ------------------------------------------------------------------------------------------------------
    def __getitem__(self, where):
        self._valid()
        if isinstance(where, str):
            if not isinstance(self._content, (Table, JaggedArray)):
                raise ValueError("no column named {0}".format(repr(where)))
            out = self.copy()
            out._content = self._content[where]
            return out
        if isinstance(where, tuple):
            if len(where) == 0:
                return self
            if len(where) == 1:
                where = where[0]
            elif len(where) == 2:
                node = self if (isinstance(where[0], slice) and where[0] == slice(None)) else self[where[0]]
                return node._getitem_inner(where[1])
            else:
                raise NotImplementedError("JaggedArray slicing in more than two dimensions is outside the GPU hot path (SURVEY.md section 8f-1)")

        if isinstance(where, JaggedArray):
            return self._getitem_jagged(where)

        if isinstance(where, (int, numpy.integer)):
            n = len(self)
            i = int(where)
            if i < 0:
                i += n
            if not 0 <= i < n:
                raise IndexError("index {0} is out of bounds for axis 0 with size {1}".format(where, n))
            if self._off64 is not None:
                pair = self._off64[i:i + 2].get()
                a, b = int(pair[0]), int(pair[1])
            else:
                a, b = int(self._general[0][i]), int(self._general[1][i])
            return self._map_content(lambda c: c[a:b])

        if isinstance(where, slice):
            start, stop, step = where.indices(len(self))
            if step != 1:
                # events start, start + step, ...: event-level fancy indexing with the index built on the device
                count = len(range(start, stop, step))
                picks = _ufunc.arange(count)
                if count:
                    picks = picks * step + start
                return self[picks]
            stop = max(start, stop)
            if self._off64 is not None:
                sub = self._off64[start:stop + 1]
                out = self.__class__.__new__(self.__class__)
                out._reset()
                out._content = self._content
                out._offsets = out._off64 = sub
                out._starts, out._stops = sub[:-1], sub[1:]
                if start == 0 and stop == len(self):
                    out._first, out._last = self._first, self._last
                    out._isvalid = True
                else:
                    ends = numpy.concatenate([sub[:1].get(), sub[-1:].get()])
                    out._first, out._last = int(ends[0]), int(ends[1])
                    out._isvalid = True
                return out
            return JaggedArray(self._starts[start:stop], self._stops[start:stop], self._content)

        # event-level selection with a flat boolean / integer array (jagged.py:599-605): new starts/stops over
        # the SAME content; like the reference the result is not dense and is compacted on first use
        if isinstance(where, (list, numpy.ndarray)):
            where = asdevice(numpy.asarray(where))
        if isinstance(where, DeviceArray):
            if self._off64 is not None:
                s64, e64 = self._off64[:-1], self._off64[1:]
            else:
                s64, e64 = self._general
            if where.dtype == numpy.bool_:
                if len(where) != len(self):
                    raise IndexError("boolean index did not match indexed array along axis 0; size of axis is {0} but size of corresponding boolean axis is {1}".format(len(self), len(where)))
                s_sel, e_sel = _ufunc.compress([s64, e64], where)
                return self._from_general(s_sel, e_sel, self._content)
            if where.dtype.kind in "iu":
                n = len(self)
                idx = where if where.dtype == numpy.int64 else where.astype(numpy.int64)
                if len(idx) and n == 0:
                    raise IndexError("index out of bounds for an empty JaggedArray")
                neg = idx < 0
                if bool(neg.any()):
                    idx = idx + neg.astype(numpy.int64) * n
                if len(idx) and (int(idx.min()) < 0 or int(idx.max()) >= n):
                    raise IndexError("index is out of bounds for axis 0 with size {0}".format(n))
                return self._from_general(_ufunc.take(s64, idx), _ufunc.take(e64, idx), self._content)
        raise NotImplementedError("this kind of JaggedArray index is outside the GPU hot path (SURVEY.md section 8f-1); no CPU fallback")

    def _getitem_inner(self, head):
        """a[:, k], a[:, i:j] and a[:, i:j:step] (jagged.py:615-707), built from device ufuncs."""
        if self._off64 is not None:
            s64, e64 = self._off64[:-1], self._off64[1:]
        else:
            s64, e64 = self._general
        counts = e64 - s64
        if isinstance(head, (int, numpy.integer)):
            k = int(head)
            local = (counts + k) if k < 0 else k
            ok = (local >= 0) & (local < counts) if k < 0 else (counts > k)
            if len(self) and not bool(ok.all()):
                raise IndexError("index {0} is out of bounds for jagged min size {1}".format(head, int(counts.min())))
            return self._map_content(lambda c: _ufunc.take(c, s64 + local))
        if isinstance(head, slice):
            if head.step not in (None, 1):
                return self._getitem_inner_step(s64, counts, head)

            def bound(v, default):
                if v is None:
                    return default
                if v >= 0:
                    return numpy.minimum(counts, v)
                return numpy.maximum(numpy.minimum(counts, counts + v), 0)

            lo = bound(head.start, 0)
            hi = bound(head.stop, counts)
            hi = numpy.maximum(lo, hi)
            return self._from_general(s64 + lo, s64 + hi, self._content)
        raise NotImplementedError("this kind of intra-event index is outside the GPU hot path (SURVEY.md section 8f-1)")

    def _getitem_inner_step(self, s64, counts, head):
        """a[:, i:j:step] (jagged.py:634-707): Python slice semantics inside every event.  The reference builds an
        (events x widest) index matrix and masks it; here the new counts come from the clamped bounds, one scan
        gives the new offsets and the source position of output element l of event i is first_i + l * step."""
        step = int(head.step)
        if step == 0:
            raise ValueError("slice step cannot be zero")
        n = len(counts)
        if step > 0:
            if head.start is None:
                lo = counts * 0
            elif head.start >= 0:
                lo = numpy.minimum(counts, head.start)
            else:
                lo = numpy.maximum(numpy.minimum(counts, counts + head.start), 0)
            if head.stop is None:
                hi = counts
            elif head.stop >= 0:
                hi = numpy.minimum(counts, head.stop)
            else:
                hi = numpy.maximum(numpy.minimum(counts, counts + head.stop), 0)
            hi = numpy.maximum(lo, hi)
            newcounts = (hi - lo + (step - 1)) // step
        else:
            if head.start is None:
                lo = counts - 1
            elif head.start >= 0:
                lo = numpy.minimum(counts - 1, head.start)
            else:
                lo = numpy.maximum(numpy.minimum(counts - 1, counts + head.start), -1)
            if head.stop is None:
                hi = counts * 0 - 1
            elif head.stop >= 0:
                hi = numpy.minimum(counts - 1, head.stop)
            else:
                hi = numpy.maximum(numpy.minimum(counts - 1, counts + head.stop), -1)
            hi = numpy.minimum(lo, hi)
            newcounts = (lo - hi + (-step - 1)) // (-step)
        newoffsets, total, _ = self._scan_counts(newcounts)
        local = empty(total, INDEXTYPE)
        if total > 0:
            call("jg_localindex", newoffsets.data_ptr(), n, _vp(local), runtime.stream())
        scaled = _ufunc.binary(numpy.multiply, DeviceArray(local, INDEXTYPE), numpy.int64(step))
        source = _ufunc.binary(numpy.add, scaled, s64 + lo, offsets=newoffsets, nevents=n, parent_side=1) if total > 0 else scaled
        return self._make(newoffsets, self._map_content(lambda c: _ufunc.take(c, source)), total)

    def _align(self, other):
        """_tojagged(self.starts, self.stops) of jagged.py:573/971-973: `other` must have my counts."""
        if len(other) != len(self):
            raise ValueError("cannot fit contents of JaggedArray into the given starts and stops arrays")
        b = other.compact()
        b._valid()
        a = self.compact()
        a._valid()
        if b._off64 is a._off64:
            return b
        if self.check_whole_valid:
            scal = runtime.scalars()
            call("jg_check_same_counts", a._off64.data_ptr(), b._off64.data_ptr(), len(a), _vp(scal, 1), runtime.stream())
            _, status, _ = runtime.read(scal)
            if status & _lib.ST_COUNTS_MISMATCH:
                raise ValueError("cannot fit contents of JaggedArray into the given starts and stops arrays")
        return b

    def _getitem_jagged(self, head):
        if isinstance(self._content, JaggedArray):
            a, roff, inner, n_inner = self._nested_parts()
            if len(head) != len(self):
                raise ValueError("jagged array used as index has a different shape {0} from the jagged array it is selecting from {1}".format(head.shape, self.shape))
            if isinstance(head._content, JaggedArray):          # jagged.py:538-539: index one level down
                h = a._align(head)
                return self._make(roff, inner[h._dense_content()], n_inner)
            h = head.compact()
            if isinstance(h._content, DeviceArray) and h._content.dtype == numpy.bool_:
                # a mask over the INNER EVENTS: event-level selection one level down + new outer offsets
                h = a._align(h)
                newoffsets, total, _ = self._scan_counts(h.sum())
                return self._make(newoffsets, inner[h._dense_content()], total)
            if isinstance(h._content, DeviceArray) and _isintegertype(h._content.dtype):
                # jagged.py:541-560 one level down: local indexes pick whole INNER EVENTS.  The checked gather kernel
                # (negative wrap, bounds) runs on the positions 0 .. n_inner-1 laid out like the outer level; the
                # absolute positions it returns select the inner events (event-level selection, general layout).
                positions = self._make(roff, DeviceArray(torch.arange(n_inner, dtype=torch.int64, device=runtime.device())), n_inner)
                picked = positions._getitem_jagged_int(h)
                return self._make(picked._off64, inner[picked._dense_content()], picked._last, picked._first)
            raise TypeError("jagged index must be boolean (mask) or integer (fancy indexing)")
        if not isinstance(head._content, DeviceArray):
            raise TypeError("jagged index must be boolean (mask) or integer (fancy indexing)")
        hdt = head._content.dtype
        if _isintegertype(hdt):
            return self._getitem_jagged_int(head)
        if hdt == numpy.bool_:
            return self._getitem_jagged_bool(head)
        raise TypeError("jagged index must be boolean (mask) or integer (fancy indexing)")

    def _getitem_jagged_bool(self, head):
        """jagged.py:562-589 -> jg_mask_select (count per tile, scan of the tile totals, compaction; the result's
        counts come out of the same sweep).  A mask that is a still-deferred `column > scalar` (lazy.LazyCompare)
        over single-column content is evaluated inside those sweeps instead (jg_compare_select)."""
        a, off, n, first, last = self._dense()
        head = a._align(head)
        mask = head._content
        mask_shift = head._first - first
        new_offsets = empty(n + 1, INDEXTYPE)
        new_counts = empty(n, INDEXTYPE)           # written by the same sweep: `.counts` of the result costs nothing
        scal = runtime.scalars()
        ws, wsb = runtime.workspace(n)
        columns = []
        fused = (isinstance(mask, _lazy.LazyCompare) and mask.deferred and isinstance(a._content, DeviceArray))
        if fused:
            key = mask._source
            # is the compared column the selected column itself (possibly seen through a view that starts at `first`)?
            same = (key.dtype == a._content.dtype and
                    key.tensor.data_ptr() + mask_shift * key.itemsize == a._content.tensor.data_ptr())
            # two staged columns must fit a tile's shared memory without shrinking it (float64 by another float64
            # column does not): the mask path is the better one there
            fused = same or (a._content.itemsize in (4, 8) and key.itemsize + a._content.itemsize <= 12)
        if fused:
            op, scalar = _lib.OP[mask._op], mask._scalar

        def select(col):
            out = empty(last - first, col.dtype)
            counts_ptr = _vp(new_counts) if not columns else ctypes.c_void_p(0)
            # (element tiles from a mean count of 7 on: a 7168-position event tile -- 4 CTAs per SM -- was measured on
            # Poisson(8 / 10 / 13): 54 - 55 % against 57 % for the element tiles, profiles/r2_exp42.log)
            long_events = self._long_events(n, last - first)
            if long_events and col.itemsize in (4, 8) and (not fused or same or key.itemsize == col.itemsize):
                ews, ewsb = runtime.elem_workspace(last - first, col.itemsize)
                # long events: flat compaction of 16 KB content tiles (one persistent launch) + rank at event boundaries
                if fused:
                    call("jg_compare_select_elem", col.data_ptr(), col.itemsize, col.data_ptr() if same else key.data_ptr(),
                         jg_code(key.dtype), op, ctypes.c_double(scalar), 0 if same else mask_shift, off.data_ptr(), n,
                         _vp(out), _vp(new_offsets), counts_ptr, _vp(scal), last - first, ews, ewsb, runtime.stream())
                else:
                    call("jg_mask_select_elem", col.data_ptr(), col.itemsize, mask.data_ptr(), mask_shift, off.data_ptr(), n,
                         _vp(out), _vp(new_offsets), counts_ptr, _vp(scal), last - first, ews, ewsb, runtime.stream())
            else:
                # the first column in two phases: count + scan, start reading K back, queue the compaction; the host
                # then waits for K alone and the next operation is queued while the compaction is still running
                ranges = SELECT_RANGES[0] and fused and same and col.itemsize == 4
                for phase in ((4,) if ranges else (3,) if SELECT_SINGLE_PASS[0] else ((1, 2) if not columns else (0,))):
                    if fused:
                        call("jg_compare_select_phase", col.data_ptr(), col.itemsize, col.data_ptr() if same else key.data_ptr(),
                             jg_code(key.dtype), op, ctypes.c_double(scalar), 0 if same else mask_shift, off.data_ptr(), n,
                             _vp(out), _vp(new_offsets), counts_ptr, _vp(scal), ws, wsb, phase, runtime.stream())
                    else:
                        call("jg_mask_select_phase", col.data_ptr(), col.itemsize, mask.data_ptr(), mask_shift, off.data_ptr(), n,
                             _vp(out), _vp(new_offsets), counts_ptr, _vp(scal), ws, wsb, phase, runtime.stream())
                    if phase == 1:
                        early.append(runtime.read_early(scal))
            columns.append(out)
            return out

        early = []
        picked = a._map_content(select)
        total, _, _ = early[0]() if early else runtime.read(scal)
        content = a._map_content(lambda t: DeviceArray(t[:total], _TORCH_NP(t)), picked)
        out = self._make(DeviceArray(new_offsets, INDEXTYPE), content, total)
        out._counts = DeviceArray(new_counts, INDEXTYPE)
        return out

    def _getitem_jagged_int(self, head):
        """jagged.py:541-560 -> jg_index_gather."""
        if len(head) != len(self):
            raise ValueError("jagged array used as index has a different shape {0} from the jagged array it is selecting from {1}".format(head.shape, self.shape))
        a, off, n, first, last = self._dense()
        src = head._arg_source
        if src is not None and src[0] is a._off64 and src[1] is a._content and src[2] == _ufunc.WRITE_EPOCH[0]:
            # `head` is argmin()/argmax() of this very array: the selected values were produced by that pass
            return self._make(head._off64, head._arg_values, head._last)
        h, hoff, _, hfirst, hlast = head._dense()
        index = h._content
        if index.dtype not in (numpy.dtype(numpy.int64), numpy.dtype(numpy.int32)):
            index = index.astype(numpy.int64)
        scal = runtime.scalars()

        def gather(col):
            out = empty(hlast - hfirst, col.dtype)
            if hlast > hfirst:
                call("jg_index_gather", col.data_ptr(), col.itemsize, off.data_ptr(), index.data_ptr(),
                     jg_code(index.dtype), hoff.data_ptr(), n, _vp(out), _vp(scal, 1), runtime.stream())
            return DeviceArray(out, col.dtype)

        content = a._map_content(gather)
        _, status, _ = runtime.read(scal)
        if status & _lib.ST_INDEX_OOB:
            raise IndexError("jagged array used as index contains out-of-bounds values")
        return self._make(h._rebased_offsets(), content, hlast - hfirst)

    # ------------------------------------------------------------------------------------------------------
    # broadcasting and ufuncs
    # ------------------------------------------------------------------------------------------------------
    def tojagged(self, data):
        """jagged.py:840-881 -> jg_broadcast for flat[N]; scalars fill; jagged data must have the same counts."""
        a, off, n, first, last = self._dense()
        if isinstance(data, JaggedArray):
            if len(data) != len(self):
                raise ValueError("cannot broadcast JaggedArray to match JaggedArray with a different counts")
            try:
                b = a._align(data)
            except ValueError:
                raise ValueError("cannot broadcast JaggedArray to match JaggedArray with a different counts")
            return self._make(a._rebased_offsets(), b._dense_content(), last - first)
        if _ufunc.is_scalar(data) or (hasattr(data, "__len__") and len(data) == 1 and not isinstance(data, (str, bytes))):
            value = numpy.asarray(data.get() if isinstance(data, DeviceArray) else data).reshape(-1)[0]
            np_dtype = numpy.asarray(value).dtype
            fill = torch.full((last - first,), value.item(), dtype=torch_dtype(np_dtype),
                              device=runtime.device())
            return self._make(a._rebased_offsets(), DeviceArray(fill, np_dtype), last - first)
        flat = asdevice(data)
        if len(flat) < n:
            # jagged.py:867-869 indexes data with the parents: an array that is too short fails there
            raise IndexError("index {0} is out of bounds for axis 0 with size {1}".format(n - 1, len(flat)))
        if len(flat) > n:
            flat = flat[:n]                     # ... and one that is longer is simply not read beyond len(self)
        out = empty(last - first, flat.dtype)
        if last > first:
            if self._long_events(n, last - first, 2):
                ws, wsb = runtime.elem_workspace(last - first, 4)
                call("jg_expand_elem", 2, flat.data_ptr(), flat.itemsize, off.data_ptr(), n, _vp(out), last - first, ws,
                     wsb, runtime.stream())
            else:
                call("jg_broadcast", flat.data_ptr(), flat.itemsize, off.data_ptr(), n, _vp(out), runtime.stream())
        return self._make(a._rebased_offsets(), DeviceArray(out, flat.dtype), last - first)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        """jagged.py:944-1051."""
        if "out" in kwargs:
            raise NotImplementedError("in-place operations not supported")
        if method != "__call__":
            return NotImplemented
        lead = None
        for x in inputs:
            if isinstance(x, JaggedArray):
                if lead is None or isinstance(x._content, JaggedArray) or not isinstance(lead._content, JaggedArray):
                    lead = x                      # the last jagged operand, a doubly jagged one if there is any
        a, off, n, first, last = lead._dense()
        if isinstance(a._content, JaggedArray):
            # jagged-of-jagged (jagged.py:1039-1051): flatten one level, let the inner arrays broadcast, re-wrap
            inner_ops = []
            for x in inputs:
                if isinstance(x, JaggedArray):
                    b = a._align(x) if x is not lead else a
                    # a singly jagged operand (e.g. deep.max()) has one value per inner event: the inner ufunc
                    # broadcasts it as a flat array (jagged.py:1039-1051 via _tojagged)
                    inner_ops.append(b._dense_content())
                elif _ufunc.is_scalar(x):
                    inner_ops.append(x)
                else:
                    flat = asdevice(x) if not isinstance(x, DeviceArray) else x
                    if flat.shape != (n,):
                        raise ValueError("cannot broadcast JaggedArray of shape {0} with array of shape {1}".format((n,), flat.shape))
                    inner_ops.append(a.tojagged(flat)._content)
            return self._make(a._rebased_offsets(), ufunc(*inner_ops), last - first)
        if isinstance(a._content, Table) or any(isinstance(x, JaggedArray) and isinstance(x._content, Table) for x in inputs):
            # records (table.py:668-738 under jagged.py:1039-1051): flatten, let the Table apply the ufunc column by
            # column (flat operands are broadcast to the content's length first), re-wrap on the shared offsets
            flat_ops = []
            for x in inputs:
                if isinstance(x, JaggedArray):
                    b = a._align(x) if x is not lead else a
                    flat_ops.append(b._dense_content())
                elif _ufunc.is_scalar(x):
                    flat_ops.append(x)
                elif isinstance(x, Table):                 # one record per event (rec.max(), ...): broadcast column by column
                    if len(x) != n:
                        raise ValueError("cannot broadcast JaggedArray of shape {0} with array of shape {1}".format((n,), (len(x),)))
                    flat_ops.append(x.map(lambda c: a.tojagged(c)._content))
                else:
                    flat = asdevice(x) if not isinstance(x, DeviceArray) else x
                    if flat.shape != (n,):
                        raise ValueError("cannot broadcast JaggedArray of shape {0} with array of shape {1}".format((n,), flat.shape))
                    flat_ops.append(a.tojagged(flat)._content)
            out = self._make(a._rebased_offsets(), ufunc(*flat_ops), last - first)
            out._counts = a._counts
            return out

        operands, parent_side = [], None
        for i, x in enumerate(inputs):
            if isinstance(x, JaggedArray):
                b = a._align(x) if x is not lead else a
                if isinstance(b._content, (Table, JaggedArray)):
                    raise NotImplementedError("ufuncs on Table / nested content are outside the GPU hot path")
                operands.append(b._dense_content())
            elif _ufunc.is_scalar(x):
                operands.append(x)
            else:
                flat = asdevice(x) if not isinstance(x, DeviceArray) else x
                if flat.shape != (n,):
                    raise ValueError("cannot broadcast JaggedArray of shape {0} with array of shape {1}".format((n,), flat.shape))
                if parent_side is not None:
                    flat = a.tojagged(flat)._content      # a second flat operand: materialise its broadcast
                else:
                    parent_side = i
                operands.append(flat)

        if len(operands) == 1:
            result = _ufunc.unary(ufunc, operands[0])
        elif len(operands) == 2:
            if parent_side is None:
                result = _ufunc.binary(ufunc, operands[0], operands[1])
            else:
                other = operands[1 - parent_side]
                if not isinstance(other, DeviceArray):
                    # flat (op) scalar with no dense operand cannot happen: the jagged input is always dense
                    raise TypeError("unexpected operand combination")
                result = _ufunc.binary(ufunc, operands[0], operands[1], offsets=off, nevents=n, parent_side=parent_side)
        else:
            raise NotImplementedError("ufunc {0} with {1} operands is not implemented on the device".format(ufunc.__name__, len(operands)))
        # fromcounts(stops - starts, result) (jagged.py:1045-1051): offsets rebased to zero, shared when possible
        out = self._make(a._rebased_offsets(), result, last - first)
        out._counts = a._counts
        return out

    # ------------------------------------------------------------------------------------------------------
    # reducers (base.py:190-215 -> jagged.py:1459-1565)
    # ------------------------------------------------------------------------------------------------------
    def _reduce_out_dtype(self, op, dt):
        if op in (_lib.ANY, _lib.ALL):
            return BOOLTYPE
        if op in (_lib.SUM, _lib.PROD):
            if dt.kind == "b" or (dt.kind == "i" and dt.itemsize < 8):
                return numpy.dtype(numpy.int64)
            if dt.kind == "u" and dt.itemsize < 8:
                return numpy.dtype(numpy.uint64)
            return dt
        if dt.kind == "b":
            return numpy.dtype(numpy.float64)         # bool min/max take the identity's type (jagged.py:1530-1531)
        return dt

    def _reduce(self, op):
        a, off, n, first, last = self._dense()
        if isinstance(a._content, JaggedArray):      # jagged.py:1463-1464: reduce the innermost level
            return self._make(a._rebased_offsets(), a._dense_content()._reduce(op), last - first)
        if isinstance(a._content, Table):
            out = Table.named(a._content.rowname)
            for name in a._content.columns:
                out[name] = a[name]._reduce(op)
            return out
        col = a._content
        odt = self._reduce_out_dtype(op, col.dtype)
        out = empty(n, odt)
        if n > 0:
            elem, hint = self._stage_hint(n, last - first, col.dtype)
            if elem:
                ws, wsb = runtime.elem_workspace(last - first, col.itemsize)
                call("jg_segreduce_elem", col.data_ptr(), jg_code(col.dtype), off.data_ptr(), n, op, _vp(out),
                     last - first, ws, wsb, runtime.stream())
            else:
                call("jg_segreduce", col.data_ptr(), jg_code(col.dtype), off.data_ptr(), n, op | hint, _vp(out), len(col),
                     runtime.stream())
        return DeviceArray(out, odt)

    def any(self):
        return self._reduce(_lib.ANY)

    def all(self):
        return self._reduce(_lib.ALL)

    def sum(self):
        return self._reduce(_lib.SUM)

    def prod(self):
        return self._reduce(_lib.PROD)

    def min(self):
        return self._reduce(_lib.MIN)

    def max(self):
        return self._reduce(_lib.MAX)

    def reduce(self, ufunc, identity):
        """base.py:190-191: the generic entry; the (ufunc, identity) pairs that have a segmented kernel."""
        known = {numpy.add: (_lib.SUM, 0), numpy.multiply: (_lib.PROD, 1),
                 numpy.minimum: (_lib.MIN, numpy.inf), numpy.maximum: (_lib.MAX, -numpy.inf)}
        if ufunc in known and identity == known[ufunc][1]:
            return self._reduce(known[ufunc][0])
        raise NotImplementedError("JaggedArray.reduce({0}, {1}) has no segmented kernel; no CPU fallback".format(
            getattr(ufunc, "__name__", ufunc), identity))

    def moment(self, n, weight=None):
        """base.py:217-222."""
        if weight is None:
            return numpy.true_divide((self ** n).sum(), self.count())
        return numpy.true_divide(((self * weight) ** n).sum(), (self * 0 + weight).sum())

    def count(self):
        """base.py:199-200 / jagged.py:1513-1518: number of non-NaN values per event."""
        a = self.compact()
        if isinstance(a._content, JaggedArray):          # jagged.py:1463-1464: the innermost level is counted
            a, off, n, first, last = self._dense()
            return self._make(a._rebased_offsets(), a._dense_content().count(), last - first)
        if isinstance(a._content, DeviceArray) and a._content.dtype.kind == "f":
            return (~numpy.isnan(a)).sum()
        return a.counts.astype(INDEXTYPE) if a.counts.dtype != INDEXTYPE else a.counts

    def count_nonzero(self):
        """base.py:202-203 / jagged.py:1520-1522."""
        return (self != 0).sum()

    def mean(self, weight=None):
        """base.py:224-229."""
        if weight is None:
            return numpy.true_divide(self.sum(), self.count())
        return numpy.true_divide((self * weight).sum(), (self * 0 + weight).sum())

    def var(self, weight=None, ddof=0):
        """base.py:231-244."""
        if weight is None:
            denom = self.count()
            one = numpy.true_divide(self.sum(), denom)
            two = numpy.true_divide((self ** 2).sum(), denom)
        else:
            denom = (self * 0 + weight).sum()
            one = numpy.true_divide((self * weight).sum(), denom)
            two = numpy.true_divide(((self * weight) ** 2).sum(), denom)
        if ddof != 0:
            return (two - one ** 2) * denom / (denom - ddof)
        return two - one ** 2

    def std(self, weight=None, ddof=0):
        """base.py:246-248."""
        return numpy.sqrt(self.var(weight=weight, ddof=ddof))

    def argmin(self):
        return self._argminmax(True)

    def argmax(self):
        return self._argminmax(False)

    def _argminmax(self, ismin):
        """jagged.py:1581-1601 -> jg_segargminmax_dense (optimistic one pass; exact two-step path if an event is all-NaN)."""
        a, off, n, first, last = self._dense()
        if isinstance(a._content, JaggedArray):      # jagged.py:1569-1570
            return self._make(a._rebased_offsets(), a._dense_content()._argminmax(ismin), last - first)
        if not isinstance(a._content, DeviceArray):
            raise ValueError("cannot compute arg{0} because content is not one-dimensional".format("min" if ismin else "max"))
        col = a._content
        out_offsets = empty(n + 1, INDEXTYPE)
        out_index = empty(n, INDEXTYPE)
        scal = runtime.scalars()
        ws, wsb = runtime.workspace(n)
        elem, hint = self._stage_hint(n, last - first, col.dtype, arg=True)
        if elem:
            # long events: element-tiled arg-reduce to best[] (a partial per tile, merged in tile order), then the scan +
            # compaction of jg_compact_nonneg -- N-sized passes, small next to the content when events are long
            best, best_value = empty(n, INDEXTYPE), empty(n, col.dtype)
            ews, ewsb = runtime.elem_workspace(last - first, col.itemsize)
            call("jg_segargminmax_elem", col.data_ptr(), jg_code(col.dtype), off.data_ptr(), n, 1 if ismin else 0, _vp(best),
                 _vp(best_value), last - first, ews, ewsb, runtime.stream())
            call("jg_compact_nonneg", _vp(best), n, _vp(out_offsets), _vp(out_index), _vp(scal), ws, wsb, runtime.stream())
            total, _, _ = runtime.read(scal)
            return self._make(DeviceArray(out_offsets, INDEXTYPE), DeviceArray(out_index[:total], INDEXTYPE), total)
        values = empty(n, col.dtype) if self.cache_arg_values else None
        tiles = a._tile_nonempty               # known when the array came from fromcounts: no counting pass
        call("jg_segargminmax_dense_tiles", col.data_ptr(), jg_code(col.dtype), off.data_ptr(), n, (1 if ismin else 0) | hint,
             _vp(tiles), _vp(out_offsets), _vp(out_index), _vp(values), _vp(scal), _vp(scal, 1), ws, wsb, runtime.stream())
        total, status, _ = runtime.read(scal)
        if status & _lib.ST_ARG_ALLNAN:
            values = None
            # a non-empty event holds only NaN: exact path (arg-reduce to best[], then scan + compaction)
            best = empty(n, INDEXTYPE)
            scal = runtime.scalars()
            call("jg_segargminmax", col.data_ptr(), jg_code(col.dtype), off.data_ptr(), n, 1 if ismin else 0, _vp(best),
                 ctypes.c_void_p(0), runtime.stream())
            call("jg_compact_nonneg", _vp(best), n, _vp(out_offsets), _vp(out_index), _vp(scal), ws, wsb, runtime.stream())
            total, _, _ = runtime.read(scal)
        out = self._make(DeviceArray(out_offsets, INDEXTYPE), DeviceArray(out_index[:total], INDEXTYPE), total)
        if values is not None:
            # the extreme values came out of the same pass: a[a.argmax()] on THIS array needs no gather
            out._arg_source = (a._off64, col, _ufunc.WRITE_EPOCH[0])
            out._arg_values = DeviceArray(values[:total], col.dtype)
        return out

    # ------------------------------------------------------------------------------------------------------
    # flatten
    # ------------------------------------------------------------------------------------------------------
    def flatten(self, axis=0):
        """jagged.py:1403-1429: dense layout -> a view of content (zero bytes moved)."""
        if not isinstance(axis, (numbers.Integral, numpy.integer)) or axis < 0:
            raise TypeError("axis must be a non-negative integer (can't count from the end)")
        if axis == 1:
            # jagged.py:1410-1416: merge the two innermost levels
            if not isinstance(self._content, JaggedArray):
                raise ValueError("cannot flatten an array containing scalar values")
            a, roff, inner, n_inner = self._nested_parts()
            inner = inner.compact()
            counts = self._make(roff, inner.counts, n_inner).sum()
            newoffsets, total, _ = self._scan_counts(counts)
            return self._make(newoffsets, inner.flatten(), total)
        if axis != 0:
            raise NotImplementedError("flatten(axis>1) is outside the GPU hot path")
        if isinstance(self._content, JaggedArray):
            # jagged-of-jagged: the inner array restricted to the reachable inner events (jagged.py:1421-1423)
            if len(self) == 0:
                return self._content[0:0]
            a, off, n, first, last = self._dense()
            return a._dense_content()
        if len(self) == 0:
            return self._map_content(lambda c: c[0:0])
        a, off, n, first, last = self._dense()
        return a._map_content(lambda c: c[first:last])

    # ------------------------------------------------------------------------------------------------------
    # combinatorics (jagged.py:1070-1367)
    # ------------------------------------------------------------------------------------------------------
    def _comb_offsets(self, kind, k, off_a, off_b, n):
        out_offsets = empty(n + 1, INDEXTYPE)
        scal = runtime.scalars()
        ws, wsb = runtime.workspace(n)
        call("jg_comb_offsets", kind, k, off_a.data_ptr(), off_b.data_ptr() if off_b is not None else ctypes.c_void_p(0),
             n, _vp(out_offsets), _vp(scal), ws, wsb, runtime.stream())
        total, _, _ = runtime.read(scal)
        return DeviceArray(out_offsets, INDEXTYPE), total

    def _comb_indices(self, kind, k, other=None, absolute=False):
        a, off, n, first, last = self._dense()
        off_b = None
        if kind == _lib.COMB_CROSS:
            b, off_b, nb, _, _ = other._dense()
        poff, total = self._comb_offsets(kind, k, off, off_b, n)
        ncols = k if kind == _lib.COMB_CHOOSE else 2
        cols = [empty(total, INDEXTYPE) for _ in range(ncols)]
        if total > 0:
            arr = (ctypes.c_void_p * ncols)(*[c.data_ptr() for c in cols])
            call("jg_comb_fill", _lib.comb_kind(kind, total, n), k, off.data_ptr(), off_b.data_ptr() if off_b is not None else ctypes.c_void_p(0), n,
                 poff.data_ptr(), arr, 1 if absolute else 0, runtime.stream())
        table = Table.named("tuple", *[DeviceArray(c, INDEXTYPE) for c in cols])
        return self._make(poff, table, total)

    def _comb_values(self, kind, other=None):
        """Value forms (jagged.py:1232-1367).  The columns of the result are deferred (lazy.py): a consumer that
        needs them gets all of them from one gather sweep, a ufunc chain on them is evaluated by one fused kernel
        straight from the two collections without the gathered temporaries ever being written."""
        a, off, n, first, last = self._dense()
        if kind == _lib.COMB_CROSS:
            b, off_b, nb, _, _ = other._dense()
        else:
            b, off_b = a, None
        poff, total = self._comb_offsets(kind, 2, off, off_b, n)
        ctx = _lazy.CombContext(kind, off, off_b, n, poff, total)

        picks = {}

        def side(content, which):
            """Flat columns are deferred gathers; a jagged column (combinations of whole inner lists) is picked
            event-wise with the absolute positions of the combination's items (jagged.py:1294 on a JaggedArray)."""
            leaves = list(_leaves(content))
            out = []
            for c in leaves:
                if isinstance(c, JaggedArray):
                    if not picks:
                        args = self._comb_indices(kind, 2, other=other, absolute=True)
                        picks["L"], picks["R"] = args._content["0"], args._content["1"]
                    out.append(c[picks[which]])
                elif isinstance(c, DeviceArray):
                    out.append(ctx.leaf(which, c))
                else:
                    raise NotImplementedError("combinatorics over masked content are outside the GPU hot path; call fillna() first")
            return _rebuild(content, iter(out))

        left, right = side(a._content, "L"), side(b._content, "R")
        return self._make(poff, Table.named("tuple", left, right).flattentuple(), total)      # jagged.py:1271,1294,1350

    def _nest(self, flat_result, kind, other=None):
        """nested=True (jagged.py:1255,1275,1285,1298,1335,1356): regroup the per-event combinations by their
        first element -- outer counts = the event's item count, inner counts from localindex / the other counts."""
        a = self.compact()
        if kind == _lib.COMB_CROSS:
            outer = a.counts
            inner = a.tojagged(other.counts)._content                  # every item meets all of other's items
        elif kind == _lib.COMB_PAIRS:
            outer = a.counts
            inner = (a.tojagged(a.counts) - a.localindex)._content      # item i pairs with items i .. n-1
        else:                                                           # distincts: item i pairs with i+1 .. n-1
            outer = numpy.maximum(a.counts - 1, 0)
            inner = (a.tojagged(a.counts) - a.localindex - 1)[:, :-1].flatten()
        inner_offsets, inner_total, _ = self._scan_counts(inner)
        outer_offsets, outer_total, _ = self._scan_counts(outer)
        inner_array = self._make(inner_offsets, flat_result._content, inner_total)
        return self._make(outer_offsets, inner_array, outer_total)

    def argpairs(self, nested=False):
        out = self._comb_indices(_lib.COMB_PAIRS, 2)
        return self._nest(out, _lib.COMB_PAIRS) if nested else out

    def pairs(self, nested=False):
        out = self._comb_values(_lib.COMB_PAIRS)
        return self._nest(out, _lib.COMB_PAIRS) if nested else out

    def argdistincts(self, nested=False):
        out = self._comb_indices(_lib.COMB_DISTINCTS, 2)
        return self._nest(out, _lib.COMB_DISTINCTS) if nested else out

    def distincts(self, nested=False):
        out = self._comb_values(_lib.COMB_DISTINCTS)
        return self._nest(out, _lib.COMB_DISTINCTS) if nested else out

    def argchoose(self, n):
        """jagged.py:1223-1230: k-combinations in colexicographic order."""
        if n > 5:
            raise NotImplementedError
        if n <= 1:
            raise ValueError("Choosing 0 or 1 items is trivial")
        return self._comb_indices(_lib.COMB_CHOOSE, int(n))

    def choose(self, n):
        """jagged.py:1232-1242."""
        if n > 5:
            raise NotImplementedError
        if n <= 1:
            raise ValueError("Choosing 0 or 1 items is trivial")
        args = self._comb_indices(_lib.COMB_CHOOSE, int(n), absolute=True)
        a = self.compact()
        cols = [a._map_content(lambda c, idx=args._content[name]: _ufunc.take(c, idx)) for name in args._content.columns]
        return self._make(args._off64, Table.named("tuple", *cols).flattentuple(), args._last)      # jagged.py:1241

    def _check_cross(self, other):
        if not isinstance(other, JaggedArray):
            raise TypeError("both arrays must be JaggedArrays")
        if len(self) != len(other):
            raise ValueError("both JaggedArrays must have the same length")

    def argcross(self, other, nested=False):
        self._valid()
        self._check_cross(other)
        out = self._comb_indices(_lib.COMB_CROSS, 2, other=other)
        return self._nest(out, _lib.COMB_CROSS, other=other) if nested else out

    def cross(self, other, nested=False):
        """jagged.py:1339-1367.  A `nested=True` result remembers its flat form (`_nestedcross`); crossing it again
        crosses that flat form with `other` and regroups the combinations under the first array's items, so that
        a.cross(b, nested=True).cross(c, nested=True) is [event][a][b][c] (jagged.py:1342-1365)."""
        self._valid()
        self._check_cross(other)
        thyself = self._nestedcross if self._nestedcross is not None else self
        if isinstance(thyself._content, JaggedArray) or isinstance(other._content, JaggedArray):
            raise NotImplementedError("cross over jagged-of-jagged content is outside the GPU hot path (SURVEY.md section 8f-2)")
        out = thyself._comb_values(_lib.COMB_CROSS, other=other)
        if nested:
            old = out
            out = thyself._nest(out, _lib.COMB_CROSS, other=other)
            out._nestedcross = old
        if thyself is not self:
            # jagged.py:1359-1365: the outer level of `out` counts combinations per event of `thyself`; split it by the
            # first array's items (counts // self.counts wherever self.counts != 0; both are 0 elsewhere)
            mine = self.counts
            per_item = out.counts // numpy.maximum(mine, 1)
            inner_offsets, inner_total, _ = self._scan_counts(self.tojagged(per_item).flatten())
            outer_offsets, outer_total, _ = self._scan_counts(mine)
            old = out
            out = self._make(outer_offsets, self._make(inner_offsets, out._content, inner_total), outer_total)
            out._nestedcross = old
        return out

    # ------------------------------------------------------------------------------------------------------
    # outside the hot path: fail loudly (north_star: no CPU fallback)
    # ------------------------------------------------------------------------------------------------------
    def _notimpl(self, name):
        raise NotImplementedError("JaggedArray.{0} is outside the GPU hot path (SURVEY.md section 8a/8f); no CPU fallback".format(name))

    def __setitem__(self, where, what):
        """jagged.py:789-838: add a column (str) or assign through a jagged mask / jagged integer index, in place."""
        if isinstance(where, str):
            if not isinstance(self._content, Table):
                raise ValueError("cannot assign a column to a JaggedArray whose content is not a Table")
            self._valid()
            if self._off64 is None or self._first != 0 or self._last != len(self._content):
                raise NotImplementedError("column assignment needs a compact JaggedArray that covers its content (no CPU fallback)")
            self._content[where] = self.tojagged(what)._content
            return
        if not isinstance(where, JaggedArray):
            raise TypeError("invalid index for assigning to JaggedArray: {0}".format(where))
        if not isinstance(self._content, DeviceArray):
            raise NotImplementedError("assignment into Table / nested content is outside the GPU hot path")
        if isinstance(what, JaggedArray):
            what = what.flatten()
        if len(where) != len(self):
            raise ValueError("jagged array used as index has a different shape {0} from the jagged array it is selecting from {1}".format(where.shape, self.shape))
        hdt = where._content.dtype if isinstance(where._content, DeviceArray) else None
        if hdt is None or not (_isintegertype(hdt) or hdt == numpy.bool_):
            raise TypeError("jagged index must be boolean (mask) or integer (fancy indexing)")
        self._valid()
        # absolute positions of every element in the content buffer, as a jagged array with this array's
        # structure (jagged.py:831: localindex + starts); selecting from it with `where` reuses the checked
        # mask / index kernels (negative wrap, bounds, shape agreement), then one scatter writes the values
        positions = self.localindex + self.starts
        flat = positions[where].flatten()
        if not (isinstance(what, DeviceArray) or _ufunc.is_scalar(what)):
            what = asdevice(numpy.asarray(what))
        _ufunc.put(self._content, flat, what)       # bumps _ufunc.WRITE_EPOCH: cached argmin / argmax values are dropped

    # ------------------------------------------------------------------------------------------------------
    # pad / fillna / regular (jagged.py:1851-1900, 1053-1068)
    # ------------------------------------------------------------------------------------------------------
    def pad(self, length, maskedwhen=True, clip=False, axis=0):
        """jagged.py:1851-1900 -> one jg_pad per leaf column (no parents / localindex / index temporaries)."""
        if not isinstance(axis, (numbers.Integral, numpy.integer)) or axis < 0:
            raise TypeError("axis must be a non-negative integer (can't count from the end)")
        if isinstance(maskedwhen, numpy.ma.core.MaskedConstant):
            raise NotImplementedError("numpy.ma.masked padding is outside the GPU hot path; use maskedwhen=True/False")
        length = int(length)
        if length < 0:
            raise ValueError("pad length must be non-negative")
        if axis > 0:
            if not isinstance(self._content, JaggedArray):
                raise NotImplementedError("pad(axis > 0) needs jagged content")
            a, roff, inner, n_inner = self._nested_parts()
            return self._make(roff, inner.pad(length, maskedwhen, clip, axis - 1), n_inner)
        self._valid()
        n = len(self)
        counts = self.counts if self.counts.dtype == INDEXTYPE else self.counts.astype(INDEXTYPE)
        if self._off64 is not None:
            starts = self._off64[:-1]
        else:
            starts = self._general[0]
        if clip:
            # every event gets exactly `length` slots and keeps at most `length` items
            arange = empty(n + 1, INDEXTYPE)
            seed = asdevice(numpy.array([0, n + 1], dtype=INDEXTYPE))
            call("jg_localindex", seed.data_ptr(), 1, _vp(arange), runtime.stream())
            dst = _ufunc.binary(numpy.multiply, DeviceArray(arange, INDEXTYPE), numpy.int64(length))
            kept = _ufunc.binary(numpy.minimum, counts, numpy.int64(length))
            total = n * length
        else:
            dst, total, _ = self._scan_counts(_ufunc.binary(numpy.maximum, counts, numpy.int64(length)))
            kept = counts
        invalid = empty(total, BOOLTYPE)
        first = [True]

        def padded(col):
            out = empty(total, col.dtype)
            if n > 0 and total > 0:
                call("jg_pad", col.data_ptr(), col.itemsize, starts.data_ptr(), kept.data_ptr(), dst.data_ptr(), n,
                     _vp(out), _vp(invalid) if first[0] else ctypes.c_void_p(0), runtime.stream())
                first[0] = False
            return DeviceArray(out, col.dtype)

        content = self._map_content(padded)
        mask = DeviceArray(invalid, BOOLTYPE)
        if not maskedwhen:
            mask = numpy.logical_not(mask)

        from_empty = len(self._content) == 0          # jagged.py:1860

        def wrap(col):
            out = MaskedArray(mask, col, maskedwhen=bool(maskedwhen))
            out._appended = from_empty
            return out

        content = content.map(wrap) if isinstance(content, Table) else wrap(content)
        return self._make(dst, content, total)

    def fillna(self, value):
        """jagged.py:1049-1051 (base.py:645-654): NaN and masked slots of the content become `value`."""
        def fill(col):
            if isinstance(col, MaskedArray):
                return col.fillna(value)
            if isinstance(col, Table):
                return col.map(fill)
            if isinstance(col, JaggedArray):
                return col.fillna(value)
            return fill_masked(col, None, True, value)

        out = self.copy()
        out._content = fill(self._content)
        out._arg_source = out._arg_values = None
        return out

    def regular(self):
        """jagged.py:1053-1068: all events must have the same count; the result is an (events x count) matrix."""
        a, off, n, first, last = self._dense(masked_ok=True)
        if n == 0:
            raise NotImplementedError("regular() of an empty JaggedArray is outside the GPU hot path")
        counts = a.counts
        lo, hi = counts.min(), counts.max()
        if int(lo) != int(hi):
            raise ValueError("jagged array is not regular: different elements have different counts")
        content = a._dense_content()
        if isinstance(content, MaskedArray):
            return content.get().reshape(n, int(lo))          # host masked matrix, like the reference's
        if not isinstance(content, DeviceArray):
            raise NotImplementedError("regular() of Table / nested content is outside the GPU hot path")
        return DeviceMatrix(content, (n, int(lo)))

    def argsort(self, ascending=False):
        """jagged.py:1603-1631 -> jg_argsort (one pass; events beyond 2048 elements through jg_argsort_long)."""
        a, off, n, first, last = self._dense()
        if isinstance(a._content, JaggedArray):      # jagged.py:1605-1606
            return self._make(a._rebased_offsets(), a._dense_content().argsort(ascending), last - first)
        if not isinstance(a._content, DeviceArray):
            raise NotImplementedError("argsort of Table content is outside the GPU hot path")
        col = a._content
        total = last - first
        out = empty(total, INDEXTYPE)
        if n > 0 and total > 0:
            cap = total // (_lib.SORT_LONG + 1) + 1
            scal = runtime.scalars()
            # [0] unused, [1] status, [2] number of long events, [3:] the first few of them; a bigger list if needed
            if cap <= 4:
                long_list = scal[2:]
            else:
                long_list = torch.zeros(cap + 1, dtype=torch.int64, device=scal.device)
            call("jg_argsort", col.data_ptr(), jg_code(col.dtype), off.data_ptr(), n, 1 if ascending else 0, _vp(out),
                 _vp(scal, 1), _vp(long_list), cap, runtime.stream())
            if total > _lib.SORT_LONG:                 # otherwise no event can be long: no read-back at all
                _, status, _ = runtime.read(scal)
                if status & _lib.ST_SORT_LONG:
                    listed = long_list.cpu().numpy()
                    events = numpy.sort(listed[1:1 + int(listed[0])])
                    where = torch.as_tensor(numpy.concatenate([events, events + 1]), device=scal.device)
                    bounds = off.tensor[where].cpu().numpy().reshape(2, -1)
                    for start, stop in bounds.T:
                        count = int(stop - start)
                        wsb = int(_lib.lib.jg_argsort_long_workspace(count))
                        ws = empty(wsb // 4, numpy.int32)
                        call("jg_argsort_long", col.data_ptr(), jg_code(col.dtype), int(start), count,
                             1 if ascending else 0, _vp(out, int(start) - first), _vp(ws), wsb, runtime.stream())
        res = self._make(a._rebased_offsets(), DeviceArray(out, INDEXTYPE), total)
        res._counts = a._counts
        return res

    # ------------------------------------------------------------------------------------------------------
    # concatenate (base.py:567-605 -> jagged.py:1633-1733)
    # ------------------------------------------------------------------------------------------------------
    @_bothmethod
    def concatenate(isclassmethod, cls_or_self, arrays, axis=0):
        arrays = tuple(arrays)
        if len(arrays) < 1:              # base.py:569-570: checked before `self` is put in front (a.concatenate([]) raises)
            raise ValueError("at least one array needed to concatenate")
        if isclassmethod:
            cls = cls_or_self
        else:
            cls = type(cls_or_self)
            arrays = (cls_or_self,) + arrays
        if not all(isinstance(x, JaggedArray) for x in arrays):
            raise NotImplementedError("concatenating a JaggedArray with other array types (UnionArray) is outside the GPU hot path")
        for x in arrays:
            x.valid(exception=True)
        if axis == 0:
            return cls._concatenate_axis0(arrays)
        if axis == 1:
            return cls._concatenate_axis1(arrays)
        raise NotImplementedError("axis > 1")

    @classmethod
    def _concat_flat(cls, columns):
        """numpy.concatenate of dense leaf columns: one allocation, one device copy (with cast) per input."""
        first = columns[0]
        if isinstance(first, Table):
            if not all(isinstance(c, Table) and list(c.columns) == list(first.columns) for c in columns):
                raise ValueError("cannot concatenate Tables with different fields")
            out = Table.named(first.rowname)
            for name in first.columns:
                out._contents[name] = cls._concat_flat([c[name] for c in columns])
            return out
        if isinstance(first, JaggedArray):
            if not all(isinstance(c, JaggedArray) for c in columns):
                raise ValueError("cannot concatenate nested and flat content")
            return cls._concatenate_axis0(columns)
        if not all(isinstance(c, DeviceArray) for c in columns):
            raise ValueError("cannot concatenate flat and structured content")
        dtype = numpy.result_type(*[c.dtype for c in columns])
        out = empty(sum(len(c) for c in columns), dtype)
        at = 0
        for c in columns:
            if len(c):
                c = c if c.dtype == dtype else _ufunc.cast(c, dtype)
                out[at:at + len(c)].copy_(c.tensor)
                at += len(c)
        return DeviceArray(out, dtype)

    @classmethod
    def _concatenate_axis0(cls, arrays):
        """jagged.py:1633-1649 -> jg_offsets_rebase per input + one content copy per input and leaf column."""
        parts = [x._dense() for x in arrays]
        nevents = sum(p[2] for p in parts)
        out_offsets = empty(nevents + 1, INDEXTYPE)
        at = base = 0
        for a, off, n, first, last in parts:
            call("jg_offsets_rebase", off.data_ptr(), n + 1, base - first, _vp(out_offsets, at), runtime.stream())
            at += n
            base += last - first
        content = cls._concat_flat([p[0]._dense_content() for p in parts])
        return cls._make(DeviceArray(out_offsets, INDEXTYPE), content, base)

    @classmethod
    def _concatenate_axis1(cls, arrays):
        """jagged.py:1651-1733 -> counts sum, one scan, then per leaf column ONE jg_segment_merge over the output
        layout (2..4 inputs); a single input or more than four go input by input through jg_segment_scatter."""
        if any(len(a) != len(arrays[0]) for a in arrays):
            raise ValueError("cannot concatenate JaggedArrays of different lengths with axis=1")
        parts = [x._dense() for x in arrays]
        n = parts[0][2]
        counts = [p[0].counts for p in parts]
        before = [None] * len(parts)                  # counts of the inputs in front of input k, per event
        run = counts[0]
        for k in range(1, len(parts)):
            before[k] = run
            run = _ufunc.binary(numpy.add, run, counts[k])
        out_offsets, total, _ = cls._scan_counts(run)
        merged = 2 <= len(parts) <= 4
        slots = []
        for k in range(0 if merged else len(parts)):
            slots.append(out_offsets[:-1] if before[k] is None else _ufunc.binary(numpy.add, out_offsets[:-1], before[k]))

        def merge(columns):
            first = columns[0]
            if isinstance(first, Table):
                if not all(isinstance(c, Table) and set(c.columns) == set(first.columns) for c in columns):
                    raise ValueError("cannot concatenate Tables with different fields")
                out = Table.named(first.rowname)
                for name in first.columns:
                    out._contents[name] = merge([c[name] for c in columns])
                return out
            if not all(isinstance(c, DeviceArray) for c in columns):
                raise NotImplementedError("concatenate with axis=1 is not implemented for these types")
            nonempty = [c for c in columns if len(c)]
            if not nonempty:
                dtype = columns[0].dtype
            elif all(c.dtype == numpy.bool_ for c in columns):
                dtype = BOOLTYPE
            else:
                dtype = numpy.result_type(*[c.dtype for c in nonempty])
            out = empty(total, dtype)
            if merged and n > 0 and total > 0:
                cast = [c if c.dtype == dtype or len(c) == 0 else _ufunc.cast(c, dtype) for c in columns]
                # the columns are dense contents starting at offsets[0]: un-rebased offsets with a shifted base
                ptrs = (ctypes.c_void_p * len(parts))(*[c.data_ptr(-parts[k][3]) for k, c in enumerate(cast)])
                offs = (ctypes.c_void_p * len(parts))(*[p[1].data_ptr() for p in parts])
                call("jg_segment_merge", ptrs, offs, len(parts), dtype.itemsize, out_offsets.data_ptr(), n, _vp(out),
                     runtime.stream())
                return DeviceArray(out, dtype)
            for k, c in enumerate(columns):
                if len(c) == 0 or merged:
                    continue
                c = c if c.dtype == dtype else _ufunc.cast(c, dtype)
                a, off, _, first_k, _ = parts[k]
                # the column is dense content starting at offsets[0]: pass the un-rebased offsets with a shifted base
                call("jg_segment_scatter", c.data_ptr(-first_k), c.itemsize, off.data_ptr(),
                     slots[k].data_ptr(), n, _vp(out), runtime.stream())
            return DeviceArray(out, dtype)

        content = merge([p[0]._dense_content() for p in parts])
        return cls._make(out_offsets, content, total)


def _has_masked(content):
    if isinstance(content, MaskedArray):
        return True
    if isinstance(content, Table):
        return any(_has_masked(content[name]) for name in content.columns)
    return False


def _rebuild(template, leaves):
    if isinstance(template, Table):
        out = Table.named(template.rowname)
        for name in template.columns:
            out._contents[name] = _rebuild(template[name], leaves)
        return out
    return next(leaves)


def _leaves(content):
    if isinstance(content, Table):
        for name in content.columns:
            for leaf in _leaves(content[name]):
                yield leaf
    else:
        yield content


def _TORCH_NP(t):
    from .device import _TORCH2NP
    return _TORCH2NP[t.dtype]
