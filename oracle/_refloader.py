"""Import the UNMODIFIED reference (awkward0 0.15.5) in the build container  --  test infrastructure.

``/root/reference`` only exists in the build container, never on the GPU box, so nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this file.  It is used by
``oracle/make_golden.py`` (fixture generation) and by the container-only cross-checks in
``tests/test_oracle_vs_reference.py`` (skipped when the reference tree is absent).

The shim is the one SURVEY.md section 8(c) lists: numpy >= 2 removed ``numpy.str``/``numpy.object`` and
changed ``numpy.array(..., copy=False)`` from "avoid a copy" to "never copy".
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("AWKWARD0_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "awkward0"))


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
