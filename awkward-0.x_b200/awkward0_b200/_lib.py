"""ctypes binding of libjagged_sm100a.so (the C ABI declared in include/jagged_b200.h).

The library is mandatory: there is no CPU fallback anywhere in this package.  Importing this module on a box
where the library has not been built raises ImportError with the build command.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JAGGED_B200_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libjagged_sm100a.so"))

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libjagged_sm100a.so not found at {0}; build it with `python awkward-0.x_b200/build.py` "
        "(awkward0_b200 has no CPU fallback)".format(LIB_PATH))

lib = ctypes.CDLL(LIB_PATH)

# ---- constants (mirror include/jagged_b200.h) -----------------------------------------------------------------
JG_F32, JG_F64, JG_I32, JG_I64, JG_B8, JG_U8, JG_I8, JG_I16, JG_U16, JG_U32, JG_U64 = range(11)

ST_NEGATIVE_COUNT = 1
ST_NEGATIVE_OFFSET = 2
ST_NONMONOTONE = 4
ST_BEYOND_CONTENT = 8
ST_INDEX_OOB = 16
ST_COUNTS_MISMATCH = 32
ST_NOT_OFFSETS = 64
ST_STOPS_LT_STARTS = 128
ST_ARG_ALLNAN = 256
ST_SORT_LONG = 512
SORT_LONG = 2048             # kSortLong in csrc/jg_sort.cu

SUM, PROD, MIN, MAX, ANY, ALL, COUNT_NONZERO = range(7)

OP = dict(
    ADD=0, SUB=1, MUL=2, DIV=3, FLOORDIV=4, MOD=5, POW=6, MAXIMUM=7, MINIMUM=8, ARCTAN2=9, HYPOT=10,
    GT=16, GE=17, LT=18, LE=19, EQ=20, NE=21, AND=24, OR=25, XOR=26,
    NEG=32, ABS=33, SQRT=34, EXP=35, LOG=36, SIN=37, COS=38, TAN=39, SINH=40, COSH=41, TANH=42, SQUARE=43,
    NOT=44, ISNAN=45, CAST=46, ISFINITE=47, ARCSINH=48, FLOOR=49, CEIL=50, RINT=51, SIGN=52, LOG10=53, EXP2=54,
    ARCSIN=55, ARCCOS=56, ARCTAN=57, ISINF=58, LOG2=59, LOG1P=60, EXPM1=61, RECIPROCAL=62)

RHS_ARRAY, RHS_SCALAR, RHS_PARENT = 0, 1, 2
COMB_PAIRS, COMB_DISTINCTS, COMB_CHOOSE, COMB_CROSS = 0, 1, 2, 3
WIDE_STAGE = 0x100               # hint OR-ed into `op` / `is_min` of the event-tiled reductions: mean count 7 - 14 (JG_WIDE_STAGE)
WIDER_STAGE = 0x200              # ... 14 - 28 (JG_WIDER_STAGE)
COMB_LONG = 0x100                # hint OR-ed into `kind`: events hold more than a handful of combinations each
COMB_LONG_MEAN = 8               # ... from this many combinations per event on (512 events then overflow a 4096-entry pair table)


def comb_kind(kind, total, nevents):
    """`kind` with the long-events hint when the mean number of combinations per event says so (same result either way;
    the kernels chosen differ: include/jagged_b200.h JG_COMB_LONG)."""
    return kind | COMB_LONG if total > COMB_LONG_MEAN * max(nevents, 1) else kind

# every symbol include/jagged_b200.h declares (tests/test_cabi_symbols.py checks the .so exports them all)
_vp, _i, _i64, _sz, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_double
PROTOTYPES = {
    "jg_version": (_i, []),
    "jg_last_error": (ctypes.c_char_p, []),
    "jg_device_info": (_i, [ctypes.POINTER(_i)] * 3),
    "jg_workspace_bytes": (_sz, [_i64]),
    "jg_counts2offsets": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "jg_counts2offsets_tiles": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "jg_offsets2counts": (_i, [_vp, _i64, _vp, _vp]),
    "jg_startsstops2counts": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "jg_validate_offsets": (_i, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "jg_validate_startsstops": (_i, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "jg_offsets2parents": (_i, [_vp, _i64, _vp, _vp]),
    "jg_localindex": (_i, [_vp, _i64, _vp, _vp]),
    "jg_startsstops2parents": (_i, [_vp, _vp, _i64, _vp, _i64, _vp]),
    "jg_compact_gather": (_i, [_vp, _i, _vp, _vp, _i64, _vp, _vp]),
    "jg_segment_scatter": (_i, [_vp, _i, _vp, _vp, _i64, _vp, _vp]),
    "jg_segment_merge": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i, _i, _vp, _i64, _vp, _vp]),
    "jg_offsets_rebase": (_i, [_vp, _i64, _i64, _vp, _vp]),
    "jg_put": (_i, [_vp, _i, _vp, _i64, _vp, ctypes.c_uint64, _i, _vp]),
    "jg_unpack_bits": (_i, [_vp, _i64, _i64, _vp, _vp]),
    "jg_pad": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "jg_fill_masked": (_i, [_vp, _i, _vp, _i, ctypes.c_double, _i64, _vp, _i64, _vp]),
    "jg_argsort": (_i, [_vp, _i, _vp, _i64, _i, _vp, _vp, _vp, _i64, _vp]),
    "jg_argsort_long_workspace": (_i64, [_i64]),
    "jg_argsort_long": (_i, [_vp, _i, _i64, _i64, _i, _vp, _vp, _sz, _vp]),
    "jg_segreduce": (_i, [_vp, _i, _vp, _i64, _i, _vp, _i64, _vp]),
    "jg_segargminmax": (_i, [_vp, _i, _vp, _i64, _i, _vp, _vp, _vp]),
    "jg_segargminmax_dense": (_i, [_vp, _i, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "jg_segargminmax_dense_tiles": (_i, [_vp, _i, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "jg_compact_nonneg": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "jg_mask_select": (_i, [_vp, _i, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "jg_compare_select": (_i, [_vp, _i, _vp, _i, _i, _d, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "jg_mask_select_phase": (_i, [_vp, _i, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "jg_compare_select_phase": (_i, [_vp, _i, _vp, _i, _i, _d, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "jg_check_same_counts": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "jg_index_gather": (_i, [_vp, _i, _vp, _vp, _i, _vp, _i64, _vp, _vp, _vp]),
    "jg_flat_compact": (_i, [_vp, _i, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "jg_take": (_i, [_vp, _i, _vp, _i64, _vp, _vp]),
    "jg_broadcast": (_i, [_vp, _i, _vp, _i64, _vp, _vp]),
    "jg_binary": (_i, [_i, _vp, _i, _vp, _i, _i, _d, _i64, _i, _i, _vp, _i, _i64, _vp, _i64, _vp]),
    "jg_unary": (_i, [_i, _vp, _i, _i, _vp, _i, _i64, _vp]),
    "jg_flat_reduce": (_i, [_vp, _i, _i64, _i, _vp, _vp, _sz, _vp]),
    "jg_comb_offsets": (_i, [_i, _i, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "jg_comb_fill": (_i, [_i, _i, _vp, _vp, _i64, _vp, ctypes.POINTER(_vp), _i, _vp]),
    "jg_comb_gather_multi": (_i, [_i, _vp, _vp, _i64, _vp, _i, _i, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i,
                                  ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp]),
    "jg_comb_gather2": (_i, [_i, _vp, _vp, _i64, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "jg_comb_eval": (_i, [_i, _vp, _vp, _i64, _vp, _i, _vp, _vp, _vp]),
    "jg_elem_workspace_bytes": (_sz, [_i64, _i]),
    "jg_segreduce_elem": (_i, [_vp, _i, _vp, _i64, _i, _vp, _i64, _vp, _sz, _vp]),
    "jg_segargminmax_elem": (_i, [_vp, _i, _vp, _i64, _i, _vp, _vp, _i64, _vp, _sz, _vp]),
    "jg_mask_select_elem": (_i, [_vp, _i, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "jg_compare_select_elem": (_i, [_vp, _i, _vp, _i, _i, _d, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "jg_expand_elem": (_i, [_i, _vp, _i, _vp, _i64, _vp, _i64, _vp, _sz, _vp]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)          # AttributeError here = the .so is stale: rebuild
    _fn.restype = _res
    _fn.argtypes = _args


class JaggedLibError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        raise JaggedLibError("{0} failed (rc={1}): {2}".format(what, rc, lib.jg_last_error().decode()))


# kernels launched per entry point (everything else launches one); bench.py reports the count as gpu_launches
LAUNCHES = [0]
_KERNELS = {"jg_flat_reduce": 2, "jg_startsstops2parents": 2, "jg_validate_startsstops": 2,
            "jg_mask_select": 3, "jg_compare_select": 3, "jg_segargminmax_dense": 3, "jg_segargminmax_dense_tiles": 2,
            "jg_flat_compact": 2, "jg_segreduce_elem": 3, "jg_segargminmax_elem": 3, "jg_mask_select_elem": 5,
            "jg_compare_select_elem": 5, "jg_expand_elem": 2}


_PHASED = {"jg_mask_select_phase": (3, 2, 1, 2), "jg_compare_select_phase": (3, 2, 1, 2, 1)}   # launches of phase 0 / 1 / 2 / 3 (/ 4)


def call(name, *args):
    """Call a C entry point that returns an int status and raise on failure."""
    if name in _PHASED:
        LAUNCHES[0] += _PHASED[name][args[-2]]
    else:
        LAUNCHES[0] += _KERNELS.get(name, 1)
    check(getattr(lib, name)(*args), name)
