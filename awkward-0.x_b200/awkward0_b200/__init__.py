"""awkward0_b200 -- the awkward0 JaggedArray hot path on B200 (sm_100a): same Python API, CUDA kernels behind a
ctypes C ABI, device-resident buffers, no CPU fallback.

    from awkward0_b200 import JaggedArray
    a = JaggedArray.fromcounts(counts, pt)          # host arrays are copied to HBM once
    lead = a[a.argmax()]; sel = a[a > 20]; s = a.sum(); p = a.pairs()
"""
from ._lib import LIB_PATH, JaggedLibError  # noqa: F401  (fails loudly when the .so is missing)
from .device import DeviceArray, DeviceMatrix, asdevice  # noqa: F401
from .masked import MaskedArray  # noqa: F401
from .table import Table  # noqa: F401
from .jagged import JaggedArray  # noqa: F401
from .arrow import fromarrow, fromparquet, toarrow, toparquet  # noqa: F401
from .ingest import StreamedJaggedArray  # noqa: F401
from .persist import load, save  # noqa: F401  (.awkd files: awkward0.save / awkward0.load)


def concatenate(arrays, axis=0):
    """awkward0/__init__.py:21-22."""
    return JaggedArray.concatenate(arrays, axis=axis)


def fromiter(iterable):
    """awkward0/__init__.py:24 (generate.py): lists of (lists of ...) numbers; other element types are out of scope."""
    return JaggedArray.fromiter(iterable)


from . import lazy  # noqa: F401,E402  (deferred pair columns + fused ufunc chains; lazy.ENABLED[0] = False switches fusion off)

__version__ = "0.1.0"
__all__ = ["JaggedArray", "Table", "DeviceArray", "DeviceMatrix", "MaskedArray", "asdevice", "fromarrow", "toarrow", "StreamedJaggedArray", "save", "load", "concatenate", "fromiter", "fromparquet", "toparquet"]
