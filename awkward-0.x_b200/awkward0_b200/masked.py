"""Device MaskedArray: what JaggedArray.pad() puts in `content` (awkward0/array/masked.py:39-406, the part the
pad -> fillna -> regular idiom touches).  A byte mask plus a content column, both in HBM; the only compute is
fillna (jg_fill_masked).  Everything else about masked arrays is outside the GPU hot path."""
import ctypes

import numpy

from ._lib import call
from .device import BOOLTYPE, DeviceArray, DeviceMatrix, asdevice, empty, jg_code, runtime


class MaskedArray(object):
    def __init__(self, mask, content, maskedwhen=True):
        mask = mask if isinstance(mask, DeviceArray) else asdevice(numpy.asarray(mask, dtype=BOOLTYPE))
        content = content if isinstance(content, DeviceArray) else asdevice(numpy.asarray(content))
        if mask.dtype != BOOLTYPE:
            raise TypeError("mask must have boolean dtype")          # masked.py:95-96
        if len(mask) < len(content):
            raise ValueError("mask must not be shorter than content")
        self._mask, self._content, self._maskedwhen = mask, content, bool(maskedwhen)
        self._appended = False       # pad() of EMPTY content: fillna promotes the dtype with the value's (see fillna)

    mask = property(lambda self: self._mask)
    content = property(lambda self: self._content)
    maskedwhen = property(lambda self: self._maskedwhen)
    dtype = property(lambda self: self._content.dtype)
    shape = property(lambda self: (len(self._content),))

    def __len__(self):
        return len(self._content)

    @property
    def nbytes(self):
        return self._mask.nbytes + self._content.nbytes

    def __getitem__(self, where):
        if isinstance(where, slice):
            return MaskedArray(self._mask[where], self._content[where], self._maskedwhen)
        raise NotImplementedError("MaskedArray indexing other than slices is outside the GPU hot path; call fillna() first")

    def boolmask(self, maskedwhen=True):
        """masked.py:173-181."""
        if bool(maskedwhen) == self._maskedwhen:
            return self._mask
        return numpy.logical_not(self._mask)

    def fillna(self, value):
        """masked.py:401-406: masked slots and NaN become `value` (cast to the content dtype).

        pad() of empty content is an IndexedMaskedArray of -1s in the reference (jagged.py:1860-1864) and its fillna
        appends the value to the content (masked.py:910-920), which promotes the dtype: same here."""
        content = self._content
        if self._appended:
            promoted = numpy.result_type(content.dtype, numpy.asarray(value).dtype)
            if promoted != content.dtype:
                content = content.astype(promoted)
        return fill_masked(content, self._mask, self._maskedwhen, value)

    def get(self):
        """Host copy as a numpy.ma.MaskedArray (device -> host)."""
        return numpy.ma.MaskedArray(self._content.get(), mask=self.boolmask(True).get()[:len(self._content)])

    def tolist(self):
        return self.get().tolist()

    def regular(self):
        return self.get()

    def __repr__(self):
        return "<MaskedArray {0} at cuda>".format(self.tolist() if len(self) <= 8 else "[...]")


def fill_masked(content, mask, maskedwhen, value):
    """jg_fill_masked on one column; `mask` may be None (NaN replacement only, base.py:645-652)."""
    dt = content.dtype
    if dt.kind == "f":
        vf, vi = float(value), 0
    elif dt.kind == "b":
        vf, vi = 0.0, int(bool(value))
    else:
        vf, vi = 0.0, int(value)
    out = empty(len(content), dt)
    call("jg_fill_masked", content.data_ptr(), jg_code(dt), mask.data_ptr() if mask is not None else ctypes.c_void_p(0),
         1 if maskedwhen else 0, ctypes.c_double(vf), ctypes.c_int64(vi), ctypes.c_void_p(out.data_ptr()), len(content),
         runtime.stream())
    return DeviceArray(out, dt)


__all__ = ["MaskedArray", "DeviceMatrix", "fill_masked"]
