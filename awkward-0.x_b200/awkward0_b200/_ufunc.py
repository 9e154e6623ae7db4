"""numpy ufunc -> CUDA kernel dispatch shared by DeviceArray and JaggedArray.__array_ufunc__.

numpy decides dtypes: the result dtype is obtained by calling the ufunc itself on zero-length host dummies
(so NEP-50 weak Python scalars, true_divide of integers, comparisons -> bool ... all follow numpy exactly);
the kernels convert operands to the compute dtype on load (jg_binary / jg_unary in include/jagged_b200.h).
"""
import ctypes
import numbers

import numpy

from . import _lib
from ._lib import OP, call
from .device import DeviceArray, asdevice, empty, jg_code, runtime, torch_dtype

_BINARY = {
    numpy.add: "ADD", numpy.subtract: "SUB", numpy.multiply: "MUL", numpy.true_divide: "DIV",
    numpy.floor_divide: "FLOORDIV", numpy.remainder: "MOD", numpy.power: "POW", numpy.maximum: "MAXIMUM",
    numpy.minimum: "MINIMUM", numpy.arctan2: "ARCTAN2", numpy.hypot: "HYPOT",
    numpy.greater: "GT", numpy.greater_equal: "GE", numpy.less: "LT", numpy.less_equal: "LE",
    numpy.equal: "EQ", numpy.not_equal: "NE",
    numpy.bitwise_and: "AND", numpy.bitwise_or: "OR", numpy.bitwise_xor: "XOR",
    numpy.logical_and: "AND", numpy.logical_or: "OR", numpy.logical_xor: "XOR",
}
_LOGICAL = (numpy.logical_and, numpy.logical_or, numpy.logical_xor)
_COMPARE = ("GT", "GE", "LT", "LE", "EQ", "NE")

_UNARY = {
    numpy.negative: "NEG", numpy.absolute: "ABS", numpy.sqrt: "SQRT", numpy.exp: "EXP", numpy.log: "LOG",
    numpy.sin: "SIN", numpy.cos: "COS", numpy.tan: "TAN", numpy.sinh: "SINH", numpy.cosh: "COSH",
    numpy.tanh: "TANH", numpy.square: "SQUARE", numpy.invert: "NOT", numpy.logical_not: "NOT",
    numpy.isnan: "ISNAN", numpy.isfinite: "ISFINITE", numpy.isinf: "ISINF", numpy.arcsinh: "ARCSINH",
    numpy.floor: "FLOOR", numpy.ceil: "CEIL", numpy.rint: "RINT", numpy.sign: "SIGN", numpy.log10: "LOG10",
    numpy.exp2: "EXP2", numpy.arcsin: "ARCSIN", numpy.arccos: "ARCCOS", numpy.arctan: "ARCTAN",
    numpy.log2: "LOG2", numpy.log1p: "LOG1P", numpy.expm1: "EXPM1", numpy.reciprocal: "RECIPROCAL",
    numpy.positive: "CAST", numpy.fabs: "ABS",
}

_COMPUTE_OK = (_lib.JG_F32, _lib.JG_F64, _lib.JG_I32, _lib.JG_I64, _lib.JG_U8, _lib.JG_B8)
# Narrow integer arithmetic runs in the next kernel-native wider type and is truncated back by a cast pass:
# + - * & | ^ ~ and powers are ring homomorphisms Z/2^32 -> Z/2^16 ..., and // % comparisons of values that are
# exactly representable give the same answers.  uint64 shares bits with int64 for the wrap-around subset.
_WIDEN = {
    numpy.dtype(numpy.int8): numpy.dtype(numpy.int32), numpy.dtype(numpy.int16): numpy.dtype(numpy.int32),
    numpy.dtype(numpy.uint16): numpy.dtype(numpy.int32), numpy.dtype(numpy.uint32): numpy.dtype(numpy.int64),
}
_U64_BITWISE_SAFE = ("ADD", "SUB", "MUL", "AND", "OR", "XOR", "NEG", "NOT", "SQUARE", "CAST")


def _native_compute(name, rdt, cdt):
    """(compute dtype the kernel runs in, dtype it stores, fix-up) for numpy's (rdt, cdt)."""
    if cdt in _WIDEN:
        wide = _WIDEN[cdt]
        return wide, (wide if rdt == cdt else rdt), ("cast" if rdt == cdt else None)
    if cdt == numpy.dtype(numpy.uint64) and rdt == cdt and name in _U64_BITWISE_SAFE:
        return numpy.dtype(numpy.int64), numpy.dtype(numpy.int64), "view"
    return cdt, rdt, None


def _finish(out, kdt, rdt, fix):
    if fix == "cast":
        return cast(DeviceArray(out, kdt), rdt)
    if fix == "view":
        return DeviceArray(out.view(torch_dtype(rdt)), rdt)
    return DeviceArray(out, rdt)


def _lazy():
    from . import lazy          # lazy.py imports device.py only; imported late to keep the module graph acyclic
    if not lazy._UFUNC_OF:
        for table in (_UNARY, _BINARY):
            for uf, nm in table.items():
                lazy._UFUNC_OF.setdefault(nm, uf)
    return lazy


def is_scalar(x):
    return isinstance(x, (numbers.Number, bool, numpy.bool_, numpy.number)) or (isinstance(x, numpy.ndarray) and x.ndim == 0)


def _dummy(x):
    """Zero-length host stand-in with the dtype semantics numpy would see."""
    if isinstance(x, DeviceArray):
        return numpy.empty(0, dtype=x.dtype)
    if isinstance(x, numpy.ndarray) and x.ndim == 0:
        return x[()]
    return x


def result_dtypes(ufunc, operands):
    """(result dtype, compute dtype) numpy would use for ufunc(*operands)."""
    dummies = [_dummy(x) for x in operands]
    with numpy.errstate(all="ignore"):
        res = ufunc(*dummies)
    if isinstance(res, tuple):
        raise NotImplementedError("multi-output ufunc {0} is not supported on the device".format(ufunc.__name__))
    rdt = numpy.asarray(res).dtype
    name = (_BINARY.get(ufunc) or _UNARY.get(ufunc))
    if ufunc in _LOGICAL or (ufunc is numpy.logical_not):
        cdt = numpy.dtype(numpy.bool_)
    elif name in _COMPARE or name in ("ISNAN", "ISFINITE", "ISINF"):
        cdt = numpy.result_type(*dummies)
    else:
        cdt = rdt
    return rdt, cdt


def _as_bool(x):
    if isinstance(x, DeviceArray):
        return x if x.dtype == numpy.bool_ else cast(x, numpy.bool_)
    return bool(x)


def cast(x, dtype):
    dtype = numpy.dtype(dtype)
    out = empty(len(x), dtype)
    call("jg_unary", OP["CAST"], x.data_ptr(), jg_code(x.dtype), jg_code(dtype), ctypes.c_void_p(out.data_ptr()),
         jg_code(dtype), len(x), runtime.stream())
    return DeviceArray(out, dtype)


def take(content, index):
    index = index if index.dtype == numpy.int64 else cast(index, numpy.int64)
    out = empty(len(index), content.dtype)
    call("jg_take", content.data_ptr(), content.itemsize, index.data_ptr(), len(index),
         ctypes.c_void_p(out.data_ptr()), runtime.stream())
    return DeviceArray(out, content.dtype)


def arange(n):
    """numpy.arange(n, dtype=int64) on the device: the localindex of one event of n items."""
    out = empty(n, numpy.int64)
    if n > 0:
        seed = asdevice(numpy.array([0, n], dtype=numpy.int64))
        call("jg_localindex", seed.data_ptr(), 1, ctypes.c_void_p(out.data_ptr()), runtime.stream())
    return DeviceArray(out, numpy.int64)


def nonzero(mask):
    """numpy.nonzero(mask)[0] for a flat boolean DeviceArray."""
    return compress([arange(len(mask))], mask)[0]


# Incremented by every in-place store.  Values cached from a buffer (the extreme values an argmin / argmax pass wrote
# beside its indexes) remember the epoch they were computed in and are dropped once ANY store has happened since:
# views share storage, so tracking the written object alone would miss a store through another view.
WRITE_EPOCH = [0]


def put(target, index, values, skip_negative=False):
    """target[index] = values (DeviceArray of the same length as index, or a scalar), in place (jg_put)."""
    dt = target.dtype
    if len(index) == 0:
        return
    WRITE_EPOCH[0] += 1
    _lazy().flush_all()          # deferred values are computed from the data as it was when they were written
    if isinstance(values, DeviceArray):
        if len(values) != len(index):
            raise ValueError("shape mismatch: value array of shape ({0},) could not be broadcast to indexing result of shape ({1},)".format(len(values), len(index)))
        values = values if values.dtype == dt else cast(values, dt)
        vptr, bits = values.data_ptr(), 0
    else:
        vptr = ctypes.c_void_p(0)
        bits = int(numpy.array([values]).astype(dt).view("u%d" % dt.itemsize)[0])
    index = index if index.dtype == numpy.int64 else cast(index, numpy.int64)
    call("jg_put", target.data_ptr(), dt.itemsize, index.data_ptr(), len(index), vptr, ctypes.c_uint64(bits),
         1 if skip_negative else 0, runtime.stream())


def compress(columns, mask):
    """[col[mask] for col in columns] for a flat boolean DeviceArray mask (one scan, one scatter per column)."""
    if mask.dtype != numpy.bool_:
        raise TypeError("boolean selection needs a bool mask")
    n = len(mask)
    for col in columns:
        if len(col) != n:
            raise IndexError("boolean index did not match indexed array along axis 0; size of axis is {0} but size of corresponding boolean axis is {1}".format(len(col), n))
    positions = empty(n + 1, numpy.int64)
    scal = runtime.scalars()
    ws, wsb = runtime.workspace(n)
    outs = []
    for col in columns:
        out = empty(n, col.dtype)
        call("jg_flat_compact", col.data_ptr(), col.itemsize, mask.data_ptr(), n, ctypes.c_void_p(out.data_ptr()),
             ctypes.c_void_p(positions.data_ptr()), ctypes.c_void_p(scal.data_ptr()), ws, wsb, runtime.stream())
        outs.append(out)
    total, _, _ = runtime.read(scal)
    return [DeviceArray(o[:total], c.dtype) for o, c in zip(outs, columns)]


def unary(ufunc, x):
    name = _UNARY.get(ufunc)
    if name is None:
        raise NotImplementedError("ufunc {0} is not implemented on the device".format(ufunc.__name__))
    rdt, cdt = result_dtypes(ufunc, [x])
    if isinstance(x, _lazy().LazyArray):
        node = _lazy().try_fuse(name, [x], rdt, cdt)
        if node is not None:
            return node
    if ufunc is numpy.logical_not:
        x = _as_bool(x)
    if name == "NOT" and x.dtype == numpy.bool_:
        cdt = numpy.dtype(numpy.bool_)
    kcdt, kdt, fix = _native_compute(name, rdt, cdt)
    ccode = jg_code(kcdt)
    if ccode not in _COMPUTE_OK:
        raise NotImplementedError("{0} on dtype {1} is not implemented on the device".format(ufunc.__name__, cdt))
    out = empty(len(x), kdt)
    call("jg_unary", OP[name], x.data_ptr(), jg_code(x.dtype), ccode, ctypes.c_void_p(out.data_ptr()), jg_code(kdt),
         len(x), runtime.stream())
    return _finish(out, kdt, rdt, fix)


def binary(ufunc, lhs, rhs, offsets=None, nevents=0, parent_side=None):
    """ufunc(lhs, rhs) where each side is a DeviceArray or a scalar.

    parent_side = None: arrays have equal length.  parent_side = 0 / 1: that side is a flat per-event array
    broadcast through `offsets` (DeviceArray int64, nevents + 1 entries); the other side is dense content that
    starts at offsets[0].
    """
    name = _BINARY.get(ufunc)
    if name is None:
        raise NotImplementedError("ufunc {0} is not implemented on the device".format(ufunc.__name__))
    if ufunc in _LOGICAL:
        lhs, rhs = _as_bool(lhs), _as_bool(rhs)
    rdt, cdt = result_dtypes(ufunc, [lhs, rhs])
    if parent_side is None and (isinstance(lhs, _lazy().LazyArray) or isinstance(rhs, _lazy().LazyArray)):
        node = _lazy().try_fuse(name, [lhs, rhs], rdt, cdt)
        if node is not None:
            return node
    if parent_side is None and name in _COMPARE:
        node = _lazy().try_compare(name, lhs, rhs, cdt)
        if node is not None:
            return node
    kcdt, kdt, fix = _native_compute(name, rdt, cdt)
    ccode = jg_code(kcdt)
    if ccode not in _COMPUTE_OK and not (name in _COMPARE and ccode == _lib.JG_U64):
        raise NotImplementedError("{0} with compute dtype {1} is not implemented on the device".format(ufunc.__name__, cdt))

    l_arr, r_arr = isinstance(lhs, DeviceArray), isinstance(rhs, DeviceArray)
    if not l_arr and not r_arr:
        raise TypeError("binary(): at least one operand must be a DeviceArray")
    swap = 0
    if parent_side is not None:
        if parent_side == 0:
            lhs, rhs, swap = rhs, lhs, 1                # kernel wants (dense, flat)
        kind = _lib.RHS_PARENT
        length = len(lhs)
        scalar = 0
    elif l_arr and r_arr:
        if len(lhs) != len(rhs):
            raise ValueError("operands could not be broadcast together with shapes ({0},) ({1},)".format(len(lhs), len(rhs)))
        kind, length, scalar = _lib.RHS_ARRAY, len(lhs), 0
    else:
        if not l_arr:
            lhs, rhs, swap = rhs, lhs, 1
        kind, length, scalar = _lib.RHS_SCALAR, len(lhs), rhs
        rhs = None

    if cdt.kind == "f":
        sf, si = float(scalar), 0
    elif cdt.kind == "b":
        sf, si = 0.0, int(bool(scalar))
    else:
        sf, si = 0.0, int(scalar)
        if cdt.kind == "u" and si < 0:
            raise OverflowError("Python integer {0} out of bounds for {1}".format(si, cdt))

    out = empty(length, kdt)
    call("jg_binary", OP[name], lhs.data_ptr(), jg_code(lhs.dtype),
         rhs.data_ptr() if rhs is not None else ctypes.c_void_p(0), jg_code(rhs.dtype) if rhs is not None else 0,
         kind, ctypes.c_double(sf), ctypes.c_int64(si), swap, ccode, ctypes.c_void_p(out.data_ptr()), jg_code(kdt),
         length, offsets.data_ptr() if offsets is not None else ctypes.c_void_p(0), int(nevents), runtime.stream())
    return _finish(out, kdt, rdt, fix)


def flat_ufunc(ufunc, inputs):
    """DeviceArray.__array_ufunc__: all array operands are flat and of equal length."""
    ops = []
    for x in inputs:
        if isinstance(x, DeviceArray) or is_scalar(x):
            ops.append(x)
        else:
            ops.append(asdevice(x))
    if len(ops) == 1:
        return unary(ufunc, ops[0])
    if len(ops) == 2:
        return binary(ufunc, ops[0], ops[1])
    raise NotImplementedError("ufunc {0} with {1} operands is not implemented on the device".format(ufunc.__name__, len(ops)))


_REDUCE_OUT = {
    # numpy's whole-array reduction dtypes
    _lib.SUM: lambda dt: numpy.dtype(numpy.int64) if dt.kind in "bi" and dt.itemsize < 8 or dt.kind == "b" else (numpy.dtype(numpy.uint64) if dt.kind == "u" else dt),
    _lib.PROD: lambda dt: numpy.dtype(numpy.int64) if dt.kind in "bi" and dt.itemsize < 8 or dt.kind == "b" else (numpy.dtype(numpy.uint64) if dt.kind == "u" else dt),
    _lib.MIN: lambda dt: dt,
    _lib.MAX: lambda dt: dt,
    _lib.ANY: lambda dt: numpy.dtype(numpy.bool_),
    _lib.ALL: lambda dt: numpy.dtype(numpy.bool_),
    _lib.COUNT_NONZERO: lambda dt: numpy.dtype(numpy.int64),
}


def flat_reduce_device(x, op):
    """Whole-array reduction, result left on the device as a 1-element DeviceArray."""
    if x.dtype.kind == "u" and op in (_lib.SUM, _lib.PROD):
        odt = numpy.dtype(numpy.int64)      # kernel accumulates small unsigned in int64
    else:
        odt = _REDUCE_OUT[op](x.dtype)
    if len(x) == 0 and op in (_lib.MIN, _lib.MAX):
        raise ValueError("zero-size array to reduction operation {0} which has no identity".format("minimum" if op == _lib.MIN else "maximum"))
    out = empty(1, odt)
    ws, wsb = runtime.workspace(1 << 16)
    call("jg_flat_reduce", x.data_ptr(), jg_code(x.dtype), len(x), op, ctypes.c_void_p(out.data_ptr()), ws, wsb,
         runtime.stream())
    return DeviceArray(out, odt)


def flat_reduce(x, op):
    out = flat_reduce_device(x, op).get()[0]
    want = _REDUCE_OUT[op](x.dtype)
    return out.astype(want) if out.dtype != want else out
