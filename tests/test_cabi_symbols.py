"""not gpu: the C-ABI library loads and exports every symbol include/jagged_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "jagged_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jg_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    names = declared_symbols()
    assert len(names) >= 25 and "jg_counts2offsets" in names and "jg_mask_select" in names


def test_library_exports_every_declared_symbol():
    from awkward0_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    from awkward0_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_symbols()


def test_version_and_workspace_query_without_gpu():
    from awkward0_b200 import _lib
    assert _lib.lib.jg_version() >= 100
    a, b = _lib.lib.jg_workspace_bytes(0), _lib.lib.jg_workspace_bytes(10**9)
    assert 0 < a < b
    assert isinstance(_lib.lib.jg_last_error(), bytes)


def test_no_cpu_fallback_without_a_device():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import awkward0_b200 as ak
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ak.JaggedArray.fromcounts([1, 2], [1.0, 2.0, 3.0])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "awkward-0.x_b200")
    pattern = re.compile(r"^\s*(from|import)\s+.*oracle", re.M)
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                assert not pattern.search(open(os.path.join(base, f)).read()), (base, f)


def declared_arities():
    """name -> number of parameters of every function the header declares."""
    text = open(os.path.join(ROOT, "include", "jagged_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for name, params in re.findall(r"\b(jg_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = params.strip()
        out[name] = 0 if params in ("", "void") else len(params.split(","))
    return out


def test_ctypes_prototypes_have_the_header_arity():
    """A ctypes argtypes list that is one argument short of the C declaration does not fail: it corrupts the call."""
    from awkward0_b200 import _lib
    arities = declared_arities()
    assert sorted(arities) == declared_symbols()
    wrong = {n: (len(args), arities[n]) for n, (_, args) in _lib.PROTOTYPES.items() if len(args) != arities[n]}
    assert not wrong, wrong


def test_ctypes_prototypes_have_the_header_types():
    """Pointer / int / int64_t / size_t / double / uint64_t of every parameter, header against the ctypes table."""
    from awkward0_b200 import _lib
    text = open(os.path.join(ROOT, "include", "jagged_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)

    def kind_c(param):
        param = " ".join(param.split())
        if "*" in param:
            return "ptr"
        for c, k in (("int64_t", "i64"), ("uint64_t", "u64"), ("size_t", "size"), ("double", "f64"), ("int32_t", "i32"),
                     ("int", "i32")):
            if re.search(r"\b%s\b" % c, param):
                return k
        raise AssertionError("unknown C parameter type: " + param)

    def kind_py(t):
        if t is ctypes.c_void_p or t is ctypes.c_char_p or hasattr(t, "contents") or getattr(t, "_type_", None) not in ("i", "l", "q", "L", "Q", "d", "I"):
            return "ptr"
        return {ctypes.c_int: "i32", ctypes.c_int64: "i64", ctypes.c_uint64: "u64", ctypes.c_size_t: "size",
                ctypes.c_double: "f64"}[t]

    same_width = ({"size", "u64"},)                      # ctypes aliases c_size_t to c_ulong == c_uint64 on LP64
    bad = []
    for name, params in re.findall(r"\b(jg_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = params.strip()
        cs = [] if params in ("", "void") else [kind_c(p) for p in params.split(",")]
        ps = [kind_py(t) for t in _lib.PROTOTYPES[name][1]]
        for i, (c, p) in enumerate(zip(cs, ps)):
            if c != p and not any({c, p} <= s for s in same_width):
                bad.append((name, i, c, p))
    assert not bad, bad


def test_host_rule_for_the_count_distribution():
    """Host logic, no GPU: which kernel family a reduction takes from (events, content length, dtype) --
    JaggedArray._stage_hint: 16 KB event tile up to a mean count of 7, 32 KB (JG_WIDE_STAGE) up to 14 and 64 KB
    (JG_WIDER_STAGE) up to 28 for 4-byte contents where measured to win, element tiles elsewhere; the hints are the
    header's values."""
    import re
    import numpy
    import awkward0_b200 as ak
    from awkward0_b200 import _lib
    header = open(os.path.join(ROOT, "include", "jagged_b200.h")).read()
    assert int(re.search(r"#define JG_WIDE_STAGE (0x[0-9a-f]+)", header).group(1), 16) == _lib.WIDE_STAGE
    assert int(re.search(r"#define JG_WIDER_STAGE (0x[0-9a-f]+)", header).group(1), 16) == _lib.WIDER_STAGE
    hint = ak.JaggedArray._stage_hint
    n = 1000000
    f32, f64, u32 = numpy.float32, numpy.float64, numpy.uint32
    assert hint(n, 5 * n, f32) == (False, 0) and hint(n, 5 * n, f64) == (False, 0)
    assert hint(n, 10 * n, f32) == (False, _lib.WIDE_STAGE) and hint(n, 10 * n, f32, arg=True) == (False, _lib.WIDE_STAGE)
    assert hint(n, 10 * n, f64) == (True, 0)                                  # (its stage is 32 KB already)
    assert hint(n, 17 * n, f32) == (True, 0) and hint(n, 17 * n, f32, arg=True) == (False, _lib.WIDER_STAGE)
    assert hint(n, 24 * n, f32) == (False, _lib.WIDER_STAGE)
    assert hint(n, 40 * n, f32) == (True, 0) and hint(n, 40 * n, f32, arg=True) == (True, 0)
    assert hint(n, 40 * n, u32) == (False, _lib.WIDER_STAGE)                   # (no element-tiled kernels for this dtype)
    assert hint(100, 1000, f32) == (False, 0)                                 # (tiny arrays: nothing to choose)
    assert hint(0, 0, f32) == (False, 0)


def test_staged_tile_kernels_hold_tma_bulk_copies_for_sm100a():
    """The library's device code is sm_100a only, and the kernels DESIGN.md section 3 describes as TMA-staged really
    contain 1-D bulk copies (SASS `UBLKCP`) with their mbarriers (`SYNCS`); nothing is a tensor-core kernel.  Static
    (cuobjdump on the built .so), so it runs without a GPU; `profiles/sass_evidence.py` prints the whole table."""
    import shutil
    import subprocess
    import pytest
    from awkward0_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump is not on PATH")
    elfs = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", elfs))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    staged, cur = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            staged[cur] = [False, False]
        elif cur is not None:
            if "UBLKCP" in line:
                staged[cur][0] = True
            elif "SYNCS" in line:
                staged[cur][1] = True
    assert len(staged) > 500
    assert not re.search(r"\b(UTCHMMA|UTCMMA|HMMA|IMMA|DMMA)\b", sass)      # no contraction anywhere on this path
    for family in ("segreduce_kernel", "segarg_direct_kernel", "mask_fill_kernel", "scan_win_kernel", "scan_tma_kernel",
                   "segreduce_elem_kernel", "segarg_elem_kernel", "select_elem_fill_kernel", "argsort_kernel"):
        members = [v for k, v in staged.items() if re.search(r"\d+%s[A-Z]" % family, k)]
        assert members, family
        assert all(bulk and mbar for bulk, mbar in members), family
