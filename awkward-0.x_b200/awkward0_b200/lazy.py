"""Deferred columns of pairs() / distincts() / cross() results and fused ufunc chains on them.

The reference gathers every column of both sides of a combination as soon as ``pairs()`` returns
(jagged.py:1294, 1350: ``self._content[argpairs...]``) and then runs each ufunc of, say, the invariant-mass formula
as one numpy pass with a P-sized temporary.  On the device that is 2 x columns x P words written and read back
plus one kernel per ufunc.  Here the columns of a combination are ``LazyArray`` leaves ("the left / right item's
value of column c for every combination") and ufuncs on them build an expression DAG; nothing runs until a
consumer needs device memory (``.tensor`` / ``data_ptr()``), at which point

* an expression is compiled into a small register program and evaluated by ONE kernel (``jg_comb_eval``,
  csrc/jg_comb.cu) that generates the combination indices, loads the columns it needs straight from the two
  collections, applies the same operator functors as the eager kernels and writes only the result;
* a bare leaf is gathered, together with all its still-deferred siblings, by the existing one-sweep gather
  kernels (``jg_comb_gather2`` / ``jg_comb_gather_multi``) -- exactly what the eager path did.

Everything a LazyArray cannot express (mixed dtypes, integer columns, operators outside the table below, programs
that are too long) falls back to materialising the operands and running the eager kernel, so results never depend
on whether fusion happened: the functors are shared (csrc/jg_ops.cuh) and no step is contracted with another.
"""
import ctypes

import numpy

from . import _lib
from ._lib import OP, call
from .device import DeviceArray, empty, jg_code, runtime

MAX_INS, MAX_COLS, MAX_CONSTS, MAX_SLOTS = 48, 8, 8, 10      # JG_EVAL_MAX_* of include/jagged_b200.h
EV_LOAD_LEFT, EV_LOAD_RIGHT, EV_LOAD_POS, EV_CONST = 128, 129, 130, 131

_FUSABLE_BINARY = ("ADD", "SUB", "MUL", "DIV", "FLOORDIV", "MOD", "POW", "MAXIMUM", "MINIMUM", "ARCTAN2", "HYPOT")
_FUSABLE_COMPARE = ("GT", "GE", "LT", "LE", "EQ", "NE")
_FUSABLE_UNARY = ("NEG", "ABS", "SQRT", "EXP", "LOG", "SIN", "COS", "TAN", "SINH", "COSH", "TANH", "SQUARE", "CAST",
                  "ARCSINH", "FLOOR", "CEIL", "RINT", "SIGN", "LOG10", "EXP2", "ARCSIN", "ARCCOS", "ARCTAN", "LOG2",
                  "LOG1P", "EXPM1", "RECIPROCAL")
_FLOATS = (numpy.dtype(numpy.float32), numpy.dtype(numpy.float64))

ENABLED = [True]          # tests switch fusion off to compare the two paths


class JgEvalProgram(ctypes.Structure):
    _fields_ = [("ins", ctypes.c_uint32 * MAX_INS),
                ("left", ctypes.c_void_p * MAX_COLS),
                ("right", ctypes.c_void_p * MAX_COLS),
                ("pos", ctypes.c_void_p * MAX_COLS),
                ("consts", ctypes.c_double * MAX_CONSTS),
                ("nins", ctypes.c_int32), ("nslots", ctypes.c_int32), ("result", ctypes.c_int32),
                ("out_bool", ctypes.c_int32)]


class CombContext(object):
    """One pairs / distincts / cross call: the two collections' offsets, the output offsets and P."""

    __slots__ = ("kind", "off_a", "off_b", "n", "poff", "total", "leaves")

    def __init__(self, kind, off_a, off_b, n, poff, total):
        self.kind, self.off_a, self.off_b, self.n, self.poff, self.total = kind, off_a, off_b, n, poff, total
        self.leaves = []

    def leaf(self, side, source):
        out = LazyArray(self, side, (source,), source.dtype)
        self.leaves.append(out)
        return out

    def _ob(self):
        return self.off_b.data_ptr() if self.off_b is not None else ctypes.c_void_p(0)

    def gather_leaves(self):
        """Materialise every still-deferred leaf: one jg_comb_gather2 for a lone left/right couple, else
        jg_comb_gather_multi sweeps of up to 4 same-itemsize columns per side."""
        todo = [x for x in self.leaves if x._mat is None]
        self.leaves = []
        if not todo:
            return
        for x in todo:
            x._mat = empty(self.total, x._ldtype)
        if self.total > 0:
            left = [x for x in todo if x._kind == "L"]
            right = [x for x in todo if x._kind == "R"]
            if len(left) == 1 and len(right) == 1:
                l, r = left[0], right[0]
                call("jg_comb_gather2", self.kind, self.off_a.data_ptr(), self._ob(), self.n, self.poff.data_ptr(),
                     l._args[0].data_ptr(), l._args[0].itemsize, r._args[0].data_ptr(), r._args[0].itemsize,
                     ctypes.c_void_p(l._mat.data_ptr()), ctypes.c_void_p(r._mat.data_ptr()), runtime.stream())
            else:
                for size in sorted(set(x._ldtype.itemsize for x in todo)):
                    il = [x for x in left if x._ldtype.itemsize == size]
                    ir = [x for x in right if x._ldtype.itemsize == size]
                    while il or ir:
                        cl, cr, il, ir = il[:4], ir[:4], il[4:], ir[4:]

                        def parr(ptrs):
                            return (ctypes.c_void_p * max(len(ptrs), 1))(*ptrs)
                        call("jg_comb_gather_multi", self.kind, self.off_a.data_ptr(), self._ob(), self.n,
                             self.poff.data_ptr(), size,
                             len(cl), parr([x._args[0].tensor.data_ptr() for x in cl]), parr([x._mat.data_ptr() for x in cl]),
                             len(cr), parr([x._args[0].tensor.data_ptr() for x in cr]), parr([x._mat.data_ptr() for x in cr]),
                             runtime.stream())
        for x in todo:
            x._args = None


class LazyArray(DeviceArray):
    """A P-sized column that is computed on first use.  kind "L"/"R": args = (source column,) ; kind "op":
    args = (operator name, operand, ...) with operands LazyArray (same context) / DeviceArray (P-sized) / scalar."""

    __slots__ = ("_ctx", "_kind", "_args", "_ldtype", "_mat")

    def __init__(self, ctx, kind, args, dtype):
        self._ctx, self._kind, self._args, self._ldtype, self._mat = ctx, kind, args, numpy.dtype(dtype), None

    # -- what DeviceArray keeps in slots is computed here
    @property
    def _t(self):
        if self._mat is None:
            self._materialise()
        return self._mat

    @property
    def _dtype(self):
        return self._ldtype

    @property
    def deferred(self):
        return self._mat is None

    # -- metadata never forces the computation
    @property
    def shape(self):
        return (self._ctx.total,)

    @property
    def size(self):
        return self._ctx.total

    @property
    def nbytes(self):
        return self._ctx.total * self._ldtype.itemsize

    @property
    def itemsize(self):
        return self._ldtype.itemsize

    def __len__(self):
        return self._ctx.total

    def __getitem__(self, where):
        if isinstance(where, slice) and where.indices(len(self)) == (0, len(self), 1):
            return self
        return DeviceArray.__getitem__(self, where)

    def __repr__(self):
        if self._mat is None:
            return "<LazyArray {0} x {1}, deferred {2}>".format(self._ctx.total, self._ldtype, self._kind)
        return DeviceArray.__repr__(self)

    def _materialise(self):
        if self._kind in ("L", "R"):
            self._ctx.gather_leaves()
            return
        prog = _compile(self)
        if prog is None:                               # cannot happen for nodes built by try_fuse; be safe
            raise RuntimeError("deferred expression cannot be compiled")
        ctx = self._ctx
        out = empty(ctx.total, self._ldtype)
        if ctx.total > 0:
            _lib.LAUNCHES[0] += 1
            _lib.check(_lib.lib.jg_comb_eval(ctx.kind, ctx.off_a.data_ptr(), ctx._ob(), ctx.n, ctx.poff.data_ptr(),
                                             jg_code(prog[1]), ctypes.byref(prog[0]), ctypes.c_void_p(out.data_ptr()),
                                             runtime.stream()), "jg_comb_eval")
        self._mat = out
        self._args = None


def _operand_dtype(x):
    return x.dtype if isinstance(x, DeviceArray) else None


def try_fuse(name, operands, rdt, cdt):
    """ufunc `name` on `operands` (DeviceArray / scalar) -> a deferred LazyArray, or None to run eagerly."""
    if not ENABLED[0]:
        return None
    ctx = None
    for x in operands:
        if isinstance(x, LazyArray) and x._mat is None:
            if ctx is None:
                ctx = x._ctx
            elif x._ctx is not ctx:
                return None
    if ctx is None or ctx.total == 0:
        return None
    compare = name in _FUSABLE_COMPARE
    if not (compare or name in _FUSABLE_BINARY or name in _FUSABLE_UNARY):
        return None
    if cdt not in _FLOATS or not (rdt == cdt or (compare and rdt == numpy.bool_)):
        return None
    for x in operands:
        if isinstance(x, DeviceArray) and (x.dtype != cdt or len(x) != ctx.total):
            return None
    node = LazyArray(ctx, "op", (name,) + tuple(operands), rdt)
    if _compile(node) is None:
        # too long / too many columns or live values: cut here -- the deferred operands become P-sized columns
        for x in operands:
            if isinstance(x, LazyArray):
                x._t
        if _compile(node) is None:
            return None
    return node


def _compile(root):
    """Expression DAG -> (JgEvalProgram, compute dtype) or None when it does not fit the evaluator's limits."""
    # post-order over unique nodes; operands: ("node", LazyArray op) | ("L"/"R"/"P", DeviceArray) | ("K", float)
    order, seen, keyed = [], {}, {}

    def key_of(x):
        if isinstance(x, LazyArray) and x._mat is None:
            if x._kind == "op":
                return ("node", id(x))
            return (x._kind, x._args[0].tensor.data_ptr())
        if isinstance(x, DeviceArray):
            return ("P", x.tensor.data_ptr())
        return ("K", float(x))

    stack = [(root, False)]
    while stack:
        x, done = stack.pop()
        k = key_of(x)
        if done:
            if k not in seen:
                seen[k] = True
                order.append(k)
                keyed[k] = x
            continue
        if k in seen or k in keyed:
            continue
        keyed[k] = x
        stack.append((x, True))
        if k[0] == "node":
            for y in reversed(x._args[1:]):
                stack.append((y, False))
    if len(order) > MAX_INS:
        return None

    uses = {}
    for k in order:
        if k[0] == "node":
            for y in keyed[k]._args[1:]:
                ky = key_of(y)
                uses[ky] = uses.get(ky, 0) + 1
    root_key = key_of(root)
    uses[root_key] = uses.get(root_key, 0) + 1

    prog = JgEvalProgram()
    cols = {"L": [], "R": [], "P": []}
    consts = []
    free = list(range(MAX_SLOTS - 1, -1, -1))
    slot = {}
    nslots = 0
    nins = 0
    cdt = None
    for k in order:
        x = keyed[k]
        if k[0] == "node":
            args = x._args[1:]
            srcs = [slot[key_of(y)] for y in args]
            for y in args:
                ky = key_of(y)
                uses[ky] -= 1
            for ky in set(key_of(y) for y in args):
                if uses[ky] == 0:
                    free.append(slot[ky])
            if not free:
                return None
            d = free.pop()
            a = srcs[0]
            b = srcs[1] if len(srcs) > 1 else 0
            word = OP[x._args[0]] | (d << 8) | (a << 16) | (b << 24)
        else:
            if not free:
                return None
            d = free.pop()
            if k[0] == "K":
                if k[1] not in consts:
                    if len(consts) == MAX_CONSTS:
                        return None
                    consts.append(k[1])
                word = EV_CONST | (d << 8) | (consts.index(k[1]) << 16)
            else:
                table = cols[k[0]]
                if len(table) == MAX_COLS:
                    return None
                table.append(k[1])
                dt = x._ldtype if isinstance(x, LazyArray) else x.dtype
                if cdt is None:
                    cdt = dt
                elif dt != cdt:
                    return None
                word = {"L": EV_LOAD_LEFT, "R": EV_LOAD_RIGHT, "P": EV_LOAD_POS}[k[0]] | (d << 8) | ((len(table) - 1) << 16)
        slot[k] = d
        nslots = max(nslots, d + 1)
        prog.ins[nins] = word
        nins += 1
    if cdt is None or cdt not in _FLOATS:
        return None
    for i, p in enumerate(cols["L"]):
        prog.left[i] = p
    for i, p in enumerate(cols["R"]):
        prog.right[i] = p
    for i, p in enumerate(cols["P"]):
        prog.pos[i] = p
    for i, c in enumerate(consts):
        prog.consts[i] = c
    prog.nins, prog.nslots, prog.result = nins, nslots, slot[root_key]
    prog.out_bool = 1 if root._ldtype == numpy.bool_ else 0
    return prog, cdt
