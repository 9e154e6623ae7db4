"""A thin device Table: an ordered set of equal-length columns.

Only what the JaggedArray hot path needs from the reference's awkward0/array/table.py: the result container of
pairs/cross/distincts/choose (``Table.named("tuple", left, right)``, table.py:246-250), column access by name
(table.py:587-600), ``flattentuple`` (table.py:791-820) and the record content produced by ``JaggedArray.zip``.
Row objects, views and recarray I/O are host-side conveniences of the reference and are out of scope.
"""
from collections import OrderedDict

import numpy
import numpy.lib.mixins

from .device import DeviceArray, asdevice


class _HostColumns(OrderedDict):
    """Host copies of a Table's columns; remembers whether its rows print as tuples or as records."""
    istuple = False


_COMPARISONS = (numpy.less, numpy.less_equal, numpy.equal, numpy.not_equal, numpy.greater, numpy.greater_equal)


class Table(numpy.lib.mixins.NDArrayOperatorsMixin):
    __array_priority__ = 60          # above DeviceArray: `column (op) table` is the table's business

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        """table.py:668-738: a ufunc on records is the ufunc on every column (one launch per column; operands that
        are not Tables meet every column); a comparison is the AND of the per-column comparisons (one bool array)."""
        if "out" in kwargs:
            raise NotImplementedError("in-place operations not supported")
        if method != "__call__":
            return NotImplemented
        if any(type(x).__name__ == "JaggedArray" for x in inputs):
            return NotImplemented                 # the jagged operand flattens first (jagged.py:1039-1051), then comes here
        table = None
        for x in inputs:
            if isinstance(x, Table):
                if table is None:
                    table = x
                elif set(table._contents) != set(x._contents):
                    raise ValueError("Tables have different sets of columns")
        newcolumns = OrderedDict()
        for n in table._contents:
            newcolumns[n] = ufunc(*[x._contents[n] if isinstance(x, Table) else x for x in inputs])
        if len(newcolumns) == 0:
            raise AssertionError("ufunc on a Table without columns")
        if any(ufunc is c for c in _COMPARISONS):
            out = None
            for x in newcolumns.values():
                out = x if out is None else numpy.bitwise_and(out, x)
            return out
        if isinstance(next(iter(newcolumns.values())), tuple):
            raise NotImplementedError("ufuncs with several results on Table content are outside the GPU hot path")
        out = Table.named(table.rowname)
        for n, x in newcolumns.items():
            out._contents[n] = x
        return out

    def __init__(self, columns1=None, *columns2, **columns3):
        self._contents = OrderedDict()
        self.rowname = "Row"
        seed = []
        if isinstance(columns1, dict):
            seed.extend(columns1.items())
        elif columns1 is not None:
            seed.append(("0", columns1))
            for i, x in enumerate(columns2):
                seed.append((str(i + 1), x))
        seed.extend(columns3.items())
        for n, x in seed:
            self[n] = x

    @classmethod
    def named(cls, rowname, columns1=None, *columns2, **columns3):
        out = cls(columns1, *columns2, **columns3)
        out.rowname = rowname
        return out

    @property
    def istuple(self):
        return self.rowname == "tuple"

    @property
    def columns(self):
        return list(self._contents)

    def __len__(self):
        for x in self._contents.values():
            return len(x)
        return 0

    def __setitem__(self, where, what):
        if not isinstance(where, str):
            raise TypeError("Table columns are set by name")
        if not isinstance(what, (DeviceArray, Table)) and type(what).__name__ != "JaggedArray":
            what = asdevice(what)
        if len(self._contents) != 0 and len(what) != len(self):
            raise ValueError("column {0} has length {1}, table has length {2}".format(where, len(what), len(self)))
        self._contents[where] = what

    def __getitem__(self, where):
        if isinstance(where, str):
            try:
                return self._contents[where]
            except KeyError:
                raise ValueError("no column named {0}".format(repr(where)))
        if isinstance(where, slice):
            out = Table.named(self.rowname)
            for n, x in self._contents.items():
                out._contents[n] = x[where]
            return out
        raise NotImplementedError("device Table supports column (str) and slice indexing only")

    def __contains__(self, name):
        return name in self._contents

    def map(self, fcn):
        """A new Table with fcn applied to every (leaf) column."""
        out = Table.named(self.rowname)
        for n, x in self._contents.items():
            out._contents[n] = x.map(fcn) if isinstance(x, Table) else fcn(x)
        return out

    def flattentuple(self):
        """table.py:791-820 -- the columns of a tuple-Table that are themselves tuple-Tables are spliced in place
        (recursively), so that a.cross(b).cross(c) is a flat 3-tuple; record Tables stay as they are."""
        out = Table.named(self.rowname)
        for n, x in self._contents.items():
            out._contents[n] = x.flattentuple() if isinstance(x, Table) else x
        if not self.istuple:
            return out
        flat = []
        for x in out._contents.values():
            if isinstance(x, Table) and x.istuple:
                flat.extend(x._contents.values())
            else:
                flat.append(x)
        out._contents = OrderedDict((str(i), x) for i, x in enumerate(flat))
        return out

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

    def __getattr__(self, name):
        contents = self.__dict__.get("_contents")
        if contents is not None and name in contents:
            return contents[name]
        raise AttributeError("no column named {0}".format(repr(name)))

    def unzip(self):
        return tuple(self._contents.values())

    def _host_columns(self):
        out = _HostColumns()
        out.istuple = self.istuple
        for n, x in self._contents.items():
            out[n] = x._host_columns() if isinstance(x, Table) else x.get()
        return out

    @property
    def nbytes(self):
        return sum(x.nbytes for x in self._contents.values())
