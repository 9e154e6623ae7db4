"""Deferred columns of pairs() / distincts() / cross() results and fused ufunc chains on them.

The reference gathers every column of both sides of a combination as soon as ``pairs()`` returns
(jagged.py:1294, 1350: ``self._content[argpairs...]``) and then runs each ufunc of, say, the invariant-mass formula
as one numpy pass with a P-sized temporary.  On the device that is 2 x columns x P words written and read back
plus one kernel per ufunc.  Here the columns of a combination are ``LazyArray`` leaves ("the left / right item's
value of column c for every combination") and ufuncs on them build an expression DAG; nothing runs until a
consumer needs device memory (``.tensor`` / ``data_ptr()``), at which point

* an expression is compiled into a small accumulator program and evaluated by ONE kernel (``jg_comb_eval``,
  csrc/jg_comb.cu) that generates the combination indices, loads the columns it needs straight from the two
  collections, applies the same operator functors as the eager kernels and writes only the result;
* a bare leaf is gathered, together with all its still-deferred siblings, by the existing one-sweep gather
  kernels (``jg_comb_gather2`` / ``jg_comb_gather_multi``) -- exactly what the eager path did.

Everything a LazyArray cannot express (mixed dtypes, integer columns, operators outside the table below, programs
that are too long) falls back to materialising the operands and running the eager kernel, so results never depend
on whether fusion happened: the functors are shared (csrc/jg_ops.cuh) and no step is contracted with another.
"""
import ctypes
import weakref

import numpy

from . import _lib
from ._lib import OP, call
from .device import DeviceArray, empty, jg_code, runtime

MAX_INS, MAX_COLS, MAX_CONSTS, MAX_SLOTS = 48, 8, 8, 10      # JG_EVAL_MAX_* of include/jagged_b200.h
EV_LOAD, EV_STORE = 30, 31                                      # JG_EV_* of include/jagged_b200.h
EV_SLOT, EV_LEFT, EV_RIGHT, EV_POS, EV_CONST = 0, 1, 2, 3, 4

_FUSABLE_BINARY = ("ADD", "SUB", "MUL", "DIV", "FLOORDIV", "MOD", "POW", "MAXIMUM", "MINIMUM", "ARCTAN2", "HYPOT")
_FUSABLE_COMPARE = ("GT", "GE", "LT", "LE", "EQ", "NE")
_FUSABLE_UNARY = ("NEG", "ABS", "SQRT", "EXP", "LOG", "SIN", "COS", "TAN", "SINH", "COSH", "TANH", "SQUARE", "CAST",
                  "ARCSINH", "FLOOR", "CEIL", "RINT", "SIGN", "LOG10", "EXP2", "ARCSIN", "ARCCOS", "ARCTAN", "LOG2",
                  "LOG1P", "EXPM1", "RECIPROCAL")
_FLOATS = (numpy.dtype(numpy.float32), numpy.dtype(numpy.float64))

ENABLED = [True]          # tests switch fusion off to compare the two paths
_UNCHECKED_SIZE = 20      # expression trees up to this size are not test-compiled when they grow by one node (a
                          # program has at most 1.5 steps per tree node; the rare misfit -- more than 8 distinct
                          # columns of one side -- is caught when the value is needed, see LazyArray._materialise)
_UFUNC_OF = {}            # operator name -> numpy ufunc, filled by _ufunc on import


_DEFERRED = {}      # id -> weak reference of every live array whose value may not have been computed yet


def _register(x):
    k = id(x)
    _DEFERRED[k] = weakref.ref(x, lambda _r, k=k: _DEFERRED.pop(k, None))


def flush_all():
    """Compute every deferred value now.  Called before anything is written IN PLACE (jg_put, the store side of
    __setitem__): a deferred value must not see data that changed after the expression was written."""
    for k, r in list(_DEFERRED.items()):
        x = r()
        if x is not None and x._mat is None:
            x.tensor
        _DEFERRED.pop(k, None)           # (values registered while this runs stay registered)


class JgEvalProgram(ctypes.Structure):
    _fields_ = [("ins", ctypes.c_uint32 * MAX_INS),
                ("left", ctypes.c_void_p * MAX_COLS),
                ("right", ctypes.c_void_p * MAX_COLS),
                ("pos", ctypes.c_void_p * MAX_COLS),
                ("consts", ctypes.c_double * MAX_CONSTS),
                ("nins", ctypes.c_int32), ("nslots", ctypes.c_int32), ("out_bool", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


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
                call("jg_comb_gather2", _lib.comb_kind(self.kind, self.total, self.n), self.off_a.data_ptr(), self._ob(), self.n, self.poff.data_ptr(),
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
                        call("jg_comb_gather_multi", _lib.comb_kind(self.kind, self.total, self.n), self.off_a.data_ptr(), self._ob(), self.n,
                             self.poff.data_ptr(), size,
                             len(cl), parr([x._args[0].tensor.data_ptr() for x in cl]), parr([x._mat.data_ptr() for x in cl]),
                             len(cr), parr([x._args[0].tensor.data_ptr() for x in cr]), parr([x._mat.data_ptr() for x in cr]),
                             runtime.stream())
        for x in todo:
            x._args = None


class LazyArray(DeviceArray):
    """A P-sized column that is computed on first use.  kind "L"/"R": args = (source column,) ; kind "op":
    args = (operator name, operand, ...) with operands LazyArray (same context) / DeviceArray (P-sized) / scalar."""

    __slots__ = ("_ctx", "_kind", "_args", "_ldtype", "_mat", "_size", "__weakref__")

    def __init__(self, ctx, kind, args, dtype, size=1):
        self._ctx, self._kind, self._args, self._ldtype, self._mat = ctx, kind, args, numpy.dtype(dtype), None
        self._size = size                 # nodes of the expression TREE below (an upper bound of the program length)
        _register(self)

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
        if prog is None:
            # does not fit the evaluator after all (e.g. more than 8 distinct columns of one side): evaluate the
            # operands (each may still fuse) and run this one step as an eager kernel -- same functors, same result
            from . import _ufunc
            args = list(self._args[1:])
            for x in args:
                if isinstance(x, LazyArray):
                    x.tensor
            ufunc = _UFUNC_OF[self._args[0]]
            ENABLED[0], was = False, ENABLED[0]
            try:
                out = _ufunc.unary(ufunc, args[0]) if len(args) == 1 else _ufunc.binary(ufunc, args[0], args[1])
            finally:
                ENABLED[0] = was
            self._mat = out.tensor
            self._args = None
            return
        ctx = self._ctx
        out = empty(ctx.total, self._ldtype)
        if ctx.total > 0:
            _lib.LAUNCHES[0] += 1
            _lib.check(_lib.lib.jg_comb_eval(_lib.comb_kind(ctx.kind, ctx.total, ctx.n), ctx.off_a.data_ptr(), ctx._ob(), ctx.n, ctx.poff.data_ptr(),
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
    size = 1
    for x in operands:
        if isinstance(x, DeviceArray) and (x.dtype != cdt or len(x) != ctx.total):
            return None
        size += x._size if isinstance(x, LazyArray) and x._mat is None else 1
    node = LazyArray(ctx, "op", (name,) + tuple(operands), rdt, size)
    if size > _UNCHECKED_SIZE and _compile(node) is None:
        # too long / too many columns or live values: cut here -- the deferred operands become P-sized columns
        for x in operands:
            if isinstance(x, LazyArray):
                x._t
        node._size = 1 + len(operands)
        if _compile(node) is None:
            return None
    return node


def _compile(root):
    """Expression DAG -> (JgEvalProgram, compute dtype) or None when it does not fit the evaluator's limits.

    The evaluator is an accumulator machine (csrc/jg_comb.cu:eval_program): the value being built stays in a
    register and every step combines it with ONE operand -- a column of either collection, a P-sized column, a
    constant or a spill slot.  A chain therefore never touches a slot; a binary node whose two inputs are both
    compound evaluates the deeper one first (Sethi-Ullman), parks it in a slot and folds it back with the
    `reverse` bit when it was the left input; a compound node used more than once is parked after its first
    evaluation and re-read from its slot."""
    def key_of(x):
        if isinstance(x, LazyArray) and x._mat is None:
            if x._kind == "op":
                return ("node", id(x))
            return (x._kind, x._args[0].tensor.data_ptr())
        if isinstance(x, DeviceArray):
            return ("P", x.tensor.data_ptr())
        return ("K", float(x))

    # unique nodes, use counts, dtype check (iterative: no recursion limit issues while counting)
    uses, nodes, dtypes = {}, {}, set()
    stack = [root]
    while stack:
        x = stack.pop()
        k = key_of(x)
        uses[k] = uses.get(k, 0) + 1
        if k in nodes:
            continue
        nodes[k] = x
        if k[0] == "node":
            stack.extend(x._args[1:])
        elif k[0] != "K":
            dtypes.add(x._ldtype if isinstance(x, LazyArray) else x.dtype)
        if len(nodes) > 4 * MAX_INS:
            return None
    if len(dtypes) != 1:
        return None
    cdt = dtypes.pop()
    if cdt not in _FLOATS:
        return None

    need = {}

    def need_of(x):                      # spill slots the evaluation of x occupies at its peak
        k = key_of(x)
        if k[0] != "node":
            return 0
        if k not in need:
            args = x._args[1:]
            if len(args) == 1:
                need[k] = need_of(args[0])
            else:
                a, b = need_of(args[0]), need_of(args[1])
                compound = [key_of(y)[0] == "node" for y in args]
                need[k] = max(a, b) if not all(compound) else (max(a, b) if a != b else a + 1)
        return need[k]

    prog = JgEvalProgram()
    cols = {"L": [], "R": [], "P": []}
    consts = []
    free = list(range(MAX_SLOTS - 1, -1, -1))
    parked = {}                          # key -> slot of a shared compound node
    remaining = dict(uses)
    state = {"nins": 0, "nslots": 0}
    KIND = {"L": EV_LEFT, "R": EV_RIGHT, "P": EV_POS}

    class _TooBig(Exception):
        pass

    def emit(op, kind=0, index=0, reverse=False):
        if state["nins"] == MAX_INS:
            raise _TooBig()
        prog.ins[state["nins"]] = op | (kind << 8) | ((1 if reverse else 0) << 12) | (index << 16)
        state["nins"] += 1

    def alloc():
        if not free:
            raise _TooBig()
        s = free.pop()
        state["nslots"] = max(state["nslots"], s + 1)
        return s

    def direct(x):
        """(kind, index) when x can be an operand of a step as it is, else None."""
        k = key_of(x)
        if k[0] == "node":
            if k in parked:
                return (EV_SLOT, parked[k])
            return None
        if k[0] == "K":
            if k[1] not in consts:
                if len(consts) == MAX_CONSTS:
                    raise _TooBig()
                consts.append(k[1])
            return (EV_CONST, consts.index(k[1]))
        table = cols[k[0]]
        if k[1] not in table:
            if len(table) == MAX_COLS:
                raise _TooBig()
            table.append(k[1])
        return (KIND[k[0]], table.index(k[1]))

    def used(x):
        """One use of x has been consumed: release the slot of a parked node after its last use."""
        k = key_of(x)
        remaining[k] -= 1
        if k in parked and remaining[k] == 0:
            free.append(parked.pop(k))

    def gen(x):
        """Emit steps that leave the value of x in the accumulator."""
        d = direct(x)
        if d is not None:
            emit(EV_LOAD, d[0], d[1])
            return
        name, args = x._args[0], x._args[1:]
        if len(args) == 1:
            gen(args[0])
            used(args[0])
            emit(OP[name])
        else:
            l, r = args
            if direct(r) is not None:
                gen(l)
                used(l)
                d = direct(r)              # (again: l may have been r itself and is parked now)
                emit(OP[name], d[0], d[1])
                used(r)
            elif direct(l) is not None:
                gen(r)
                used(r)
                d = direct(l)
                emit(OP[name], d[0], d[1], reverse=True)
                used(l)
            else:
                first_is_left = need_of(l) >= need_of(r)
                first, second = (l, r) if first_is_left else (r, l)
                gen(first)
                k1 = key_of(first)
                if k1 in parked:           # shared: it already sits in a slot
                    gen(second)
                    used(second)
                    d = direct(first)
                    emit(OP[name], d[0], d[1], reverse=first_is_left)
                    used(first)
                else:
                    s = alloc()
                    emit(EV_STORE, EV_SLOT, s)
                    used(first)
                    gen(second)
                    used(second)
                    emit(OP[name], EV_SLOT, s, reverse=first_is_left)
                    free.append(s)
        k = key_of(x)
        if uses[k] > 1 and k not in parked:
            parked[k] = alloc()
            emit(EV_STORE, EV_SLOT, parked[k])

    try:
        gen(root)
    except (_TooBig, RecursionError):
        return None
    for i, p in enumerate(cols["L"]):
        prog.left[i] = p
    for i, p in enumerate(cols["R"]):
        prog.right[i] = p
    for i, p in enumerate(cols["P"]):
        prog.pos[i] = p
    for i, c in enumerate(consts):
        prog.consts[i] = c
    prog.nins, prog.nslots = state["nins"], state["nslots"]
    prog.out_bool = 1 if root._ldtype == numpy.bool_ else 0
    return prog, cdt


# ---- column OP scalar, deferred so that a[a > c] never writes the mask ---------------------------------------------
_MIRROR = {"GT": "LT", "GE": "LE", "LT": "GT", "LE": "GE"}


class LazyCompare(DeviceArray):
    """`column OP scalar` for an ordering comparison (> >= < <=) of a float32 / float64 column.  Used as a jagged
    mask (jagged.py:562-589) it is evaluated inside the selection sweeps (jg_compare_select): the M-byte mask is
    neither written nor read.  Any other use computes it with the ordinary comparison kernel first."""

    __slots__ = ("_source", "_op", "_scalar", "_mat", "__weakref__")

    def __init__(self, source, op, scalar):
        self._source, self._op, self._scalar, self._mat = source, op, float(scalar), None
        _register(self)

    @property
    def _t(self):
        if self._mat is None:
            from . import _ufunc
            ENABLED[0], was = False, ENABLED[0]
            try:
                self._mat = _ufunc.binary(_UFUNC_OF[self._op], self._source, self._source.dtype.type(self._scalar)).tensor
            finally:
                ENABLED[0] = was
            self._source = None
        return self._mat

    @property
    def _dtype(self):
        return numpy.dtype(numpy.bool_)

    @property
    def deferred(self):
        return self._mat is None

    def __len__(self):
        return len(self._source) if self._mat is None else int(self._mat.shape[0])

    @property
    def shape(self):
        return (len(self),)

    @property
    def size(self):
        return len(self)

    @property
    def nbytes(self):
        return len(self)

    @property
    def itemsize(self):
        return 1

    def __getitem__(self, where):
        if self._mat is None and isinstance(where, slice):
            start, stop, step = where.indices(len(self))
            if step == 1:
                if (start, stop) == (0, len(self)):
                    return self
                return LazyCompare(self._source[start:max(start, stop)], self._op, self._scalar)
        return DeviceArray.__getitem__(self, where)

    def __repr__(self):
        if self._mat is None:
            return "<LazyCompare {0} x bool, deferred {1} {2}>".format(len(self), self._op, self._scalar)
        return DeviceArray.__repr__(self)


def try_compare(name, lhs, rhs, cdt):
    """An ordering comparison of a float column with a scalar -> LazyCompare, else None."""
    if not ENABLED[0] or name not in _MIRROR:
        return None
    if isinstance(lhs, DeviceArray) and not isinstance(rhs, DeviceArray):
        arr, scalar, op = lhs, rhs, name
    elif isinstance(rhs, DeviceArray) and not isinstance(lhs, DeviceArray):
        arr, scalar, op = rhs, lhs, _MIRROR[name]                 # c OP x  ==  x mirror(OP) c
    else:
        return None
    if isinstance(arr, (LazyArray, LazyCompare)) or arr.dtype not in _FLOATS or cdt != arr.dtype or len(arr) == 0:
        return None
    try:
        scalar = float(scalar)
    except (TypeError, ValueError):
        return None
    return LazyCompare(arr, op, scalar)
