"""Import the UNMODIFIED reference (awkward0 0.15.5) in the build container  --  test infrastructure.

``/root/reference`` only exists in the build container, never on the GPU box; there the byte-identical copy
``oracle/_ref`` (recipe: ``oracle/fetch_ref.py``, git-ignored, travels with the gpurun snapshot) is imported instead.
Used by ``oracle/make_golden.py`` (fixture generation), the cross-checks in ``tests/test_oracle_vs_reference.py``
(skipped when neither tree exists) and ``bench.py``'s CPU legs -- never by the product package.

The shim is the one SURVEY.md section 8(c) lists: numpy >= 2 removed ``numpy.str``/``numpy.object`` and
changed ``numpy.array(..., copy=False)`` from "avoid a copy" to "never copy".
"""
import os
import sys

_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")       # oracle/fetch_ref.py puts it there


def _root():
    """The reference tree itself in the build container, else the unmodified copy that travelled with the snapshot."""
    for root in (os.environ.get("AWKWARD0_REFERENCE"), "/root/reference", _COPY):
        if root and os.path.isdir(os.path.join(root, "awkward0")):
            return root
    return None


REFERENCE_ROOT = _root() or "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "awkward0"))


def is_copy():
    return os.path.abspath(REFERENCE_ROOT) == os.path.abspath(_COPY)


def load():
    import numpy
    if not getattr(numpy.array, "_awkward0_shim", False):
        original = numpy.array

        def array(*args, **kwargs):
            if kwargs.get("copy", True) is False:
                kwargs["copy"] = None
            return original(*args, **kwargs)

        array._awkward0_shim = True
        numpy.array = array
    for name, val in (("str", str), ("object", object)):
        if name not in numpy.__dict__:
            setattr(numpy, name, val)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import awkward0
    return awkward0
