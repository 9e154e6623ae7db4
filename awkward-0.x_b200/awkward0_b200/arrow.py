"""Arrow <-> device JaggedArray (awkward0/arrow.py:231-430 fromarrow, :27-229 toarrow).

Arrow delivers exactly the hot path's inputs -- a list array IS (offsets int32/int64, values) -- so ingest is one
host->device copy per buffer and `JaggedArray.fromoffsets`; nothing is re-scanned.  Supported: list / large_list
(nested to any depth) of numeric, boolean (bit-packed, unpacked by jg_unpack_bits) or struct-of-those values, and
ChunkedArrays of them (chunks are appended with JaggedArray.concatenate).  Null VALUES become a MaskedArray content
(the reference: BitMaskedArray; `fillna` / `tolist` / `counts` work, everything else asks for fillna first); null
LISTS or structs (a BitMaskedArray around the JaggedArray itself) are outside the GPU hot path: NotImplementedError.
"""
import ctypes

import numpy

from ._lib import call
from .device import BOOLTYPE, DeviceArray, asdevice, empty, runtime
from .jagged import JaggedArray
from .masked import MaskedArray
from .table import Table


def _no_nulls(array):
    if array.null_count != 0:
        raise NotImplementedError("Arrow arrays with nulls become BitMaskedArrays in awkward0 (arrow.py:274-276), which is outside the GPU hot path")


def _unpack(bits, offset, n):
    out = empty(n, BOOLTYPE)
    if n > 0:
        packed = asdevice(numpy.frombuffer(bits, dtype=numpy.uint8))
        call("jg_unpack_bits", packed.data_ptr(), offset, n, ctypes.c_void_p(out.data_ptr()), runtime.stream())
    return DeviceArray(out, BOOLTYPE)


def _values(array):
    """A non-list Arrow array -> DeviceArray / Table of DeviceArrays; numeric or boolean values with nulls ->
    MaskedArray (the reference wraps them in a BitMaskedArray over Arrow's validity bits, arrow.py:262-266; here the
    bits are unpacked to a byte mask by jg_unpack_bits, valid = unmasked)."""
    import pyarrow
    tpe = array.type
    if array.null_count != 0 and not pyarrow.types.is_struct(tpe) and array.buffers()[0] is not None:
        valid = _unpack(array.buffers()[0], array.offset, len(array))
        plain = _values_plain(array)
        return MaskedArray(valid, plain, maskedwhen=False)
    return _values_plain(array)


def _values_plain(array):
    import pyarrow
    tpe = array.type
    if pyarrow.types.is_struct(tpe):
        _no_nulls(array)
        out = Table.named("Row")
        for i in range(tpe.num_fields):
            out[tpe.field(i).name] = _content(array.field(i))
        return out
    if pyarrow.types.is_boolean(tpe):
        return _unpack(array.buffers()[1], array.offset, len(array))
    if pyarrow.types.is_integer(tpe) or pyarrow.types.is_floating(tpe):
        dtype = numpy.dtype(tpe.to_pandas_dtype())
        if dtype == numpy.dtype(numpy.float16):
            raise NotImplementedError("float16 content is not supported on the device")
        buf = array.buffers()[1]
        host = numpy.frombuffer(buf, dtype=dtype)[array.offset:array.offset + len(array)] if buf is not None else numpy.zeros(0, dtype)
        return asdevice(host)
    raise NotImplementedError("Arrow type {0} is outside the GPU hot path".format(tpe))


def _content(array):
    import pyarrow
    tpe = array.type
    if pyarrow.types.is_list(tpe) or pyarrow.types.is_large_list(tpe):
        _no_nulls(array)
        # .offsets honours the slice offset of the array; .values is the whole child (offsets index into it)
        offsets = numpy.asarray(array.offsets)
        return JaggedArray.fromoffsets(offsets, _content(array.values))
    return _values(array)


def fromarrow(obj):
    """pyarrow ListArray / LargeListArray / ChunkedArray of them -> JaggedArray on the device."""
    import pyarrow
    if isinstance(obj, pyarrow.ChunkedArray):
        chunks = [x for x in obj.chunks if len(x) > 0]
        if len(chunks) == 0:
            chunks = list(obj.chunks[:1])
        parts = [fromarrow(x) for x in chunks]
        return parts[0] if len(parts) == 1 else JaggedArray.concatenate(parts)
    if isinstance(obj, pyarrow.Array):
        out = _content(obj)
        if not isinstance(out, JaggedArray):
            raise NotImplementedError("fromarrow on the device expects a list-typed array (the JaggedArray hot path)")
        return out
    raise NotImplementedError("fromarrow of {0} is outside the GPU hot path".format(type(obj)))


def toarrow(array):
    """JaggedArray -> pyarrow.ListArray (int32 offsets when they fit, like arrow.py:128-136; large_list otherwise)."""
    import pyarrow

    def recurse(x):
        if isinstance(x, JaggedArray):
            x = x.compact()
            offsets = x._rebased_offsets().get()
            child = recurse(x._dense_content())
            if len(offsets) and offsets[-1] <= numpy.iinfo(numpy.int32).max:
                return pyarrow.ListArray.from_arrays(pyarrow.array(offsets.astype(numpy.int32)), child)
            return pyarrow.LargeListArray.from_arrays(pyarrow.array(offsets), child)
        if isinstance(x, Table):
            names = list(x.columns)
            return pyarrow.StructArray.from_arrays([recurse(x[n]) for n in names], names)
        if isinstance(x, DeviceArray):
            return pyarrow.array(x.get())
        raise NotImplementedError("toarrow of {0} is outside the GPU hot path".format(type(x)))

    if not isinstance(array, JaggedArray):
        raise NotImplementedError("toarrow on the device expects a JaggedArray")
    return recurse(array)


def toparquet(where, obj, **options):
    """arrow.py:452-523: a JaggedArray becomes a Parquet file with ONE column named "" (the reference's layout); an
    iterable of JaggedArrays writes one row group per item.  Host work is pyarrow's; the device side is the D2H
    of offsets + content in `toarrow`."""
    import pyarrow
    import pyarrow.parquet

    def convert(item, message):
        if not isinstance(item, JaggedArray):
            raise TypeError(message)
        return pyarrow.Table.from_batches([pyarrow.RecordBatch.from_arrays([toarrow(item)], [""])])

    if isinstance(obj, JaggedArray):
        items = iter([obj])
    else:
        try:
            items = iter(obj)
        except TypeError:
            raise TypeError("cannot write {0} to Parquet file".format(type(obj)))
    try:
        first = next(items)
    except StopIteration:
        raise ValueError("iterable is empty")
    table = convert(first, "cannot write iterator of {0} to Parquet file".format(type(first)))
    options["where"] = where
    if "schema" not in options:
        options["schema"] = table.schema
    writer = pyarrow.parquet.ParquetWriter(**options)
    try:
        writer.write_table(table)
        for item in items:
            writer.write_table(convert(item, "cannot write iterator of {0} to Parquet file".format(type(item))))
    finally:
        writer.close()


def fromparquet(file, columns=None):
    """arrow.py:557-578.  The reference returns a lazy ChunkedArray of VirtualArrays, one chunk per row group
    (outside the GPU hot path); here every row group is decoded by pyarrow on the host, its (offsets, values) buffers
    are copied to HBM and the groups are appended with `JaggedArray.concatenate`: the result is ONE resident
    JaggedArray for a file whose only column is "" (what `toparquet` writes), else an ordered dict column -> array."""
    import collections
    import pyarrow.parquet

    pf = pyarrow.parquet.ParquetFile(file)
    names = [n for n in pf.schema_arrow.names if columns is None or n in columns]
    parts = collections.OrderedDict((n, []) for n in names)
    for g in range(pf.num_row_groups):
        if pf.metadata.row_group(g).num_rows == 0:
            continue
        table = pf.read_row_group(g, columns=names)
        for n in names:
            parts[n].append(fromarrow(table.column(n)))
    out = collections.OrderedDict()
    for n in names:
        if not parts[n]:                               # no rows at all: an empty array of the column's type
            parts[n].append(fromarrow(pf.schema_arrow.empty_table().column(n).combine_chunks()))
        out[n] = parts[n][0] if len(parts[n]) == 1 else JaggedArray.concatenate(parts[n])
    if names == [""]:
        return out[""]
    return out
