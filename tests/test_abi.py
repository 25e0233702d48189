"""-m "not gpu": the C-ABI library builds, loads and exports every symbol include/pwgs_b200.h declares; the ctypes table
covers exactly that set; and without a CUDA device the product fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pwgs_b200.h")


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(pwgs_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from phylowgs_b200 import build, _lib
    build.build_library()
    return _lib.load()


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for must in ("pwgs_create", "pwgs_destroy", "pwgs_set_tree", "pwgs_mh_run", "pwgs_llh_rows", "pwgs_total_llh", "pwgs_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), "libpwgs_b200.so does not export %s" % name


def test_ctypes_table_matches_header(lib):
    from phylowgs_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_oracle_or_torch_in_the_product():
    """The product path must not import the oracle (or torch): it is ctypes + numpy over the CUDA library only."""
    pkg = os.path.join(ROOT, "phylowgs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            src = open(os.path.join(dirpath, f), errors="ignore").read() if f.endswith((".py", ".cu", ".cuh", ".h")) else ""
            if f.endswith(".py"):
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert not re.search(r"^\s*(from|import)\s+torch\b", src, flags=re.M), f
                assert "liboracle" not in src and "oracle/_ref" not in src and "libpwgs_ref" not in src, f
            else:
                assert not re.search(r'#include\s+"[^"]*oracle', src), f


def test_fails_loudly_without_a_gpu(lib):
    import phylowgs_b200 as P
    if P.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    a = np.ones((3, 1), dtype=np.int32)
    d = 2 * a
    with pytest.raises(P.PwgsError) as e:
        P.Engine(a, d, [0.999] * 3, [0.5] * 3, 3)
    assert "error -2" in str(e.value)          # PWGS_ECUDA: no device, and no CPU fallback exists


def test_argument_validation_happens_before_any_device_work(lib):
    from phylowgs_b200 import _lib
    h = C.c_void_p()
    a = np.array([[5]], dtype=np.int32)
    d = np.array([[3]], dtype=np.int32)          # a > d is invalid
    mu = np.array([0.5])
    rc = lib.pwgs_create(0, 1, 0, 1, _lib.p_i32(a), _lib.p_i32(d), _lib.p_f64(mu), _lib.p_f64(mu), None, None, None, None, C.byref(h))
    assert rc == -1 and b"a <= d" in lib.pwgs_last_error()
    rc = lib.pwgs_create(0, 0, 0, 1, None, None, None, None, None, None, None, None, C.byref(h))
    assert rc == -1


def test_pickle_join_gathers_and_validates(lib):
    """pwgs_pickle_join (host code): fragment | reference | tail per item, fragments alone for negative references; capacity and
    index errors are reported, nothing is written past `cap`."""
    frags = [b"alpha", b"", b"bc", b"delta!"]
    blob = b"".join(frags)
    off = np.zeros(5, dtype=np.int64)
    np.cumsum([len(f) for f in frags], out=off[1:])
    refs = np.zeros((2, 8), dtype=np.uint8)
    for k, r in enumerate([b"h\x07", b"j\x01\x02\x03\x04"]):
        refs[k, 0] = len(r)
        refs[k, 1:1 + len(r)] = np.frombuffer(r, dtype=np.uint8)
    ref_of = np.array([1, 0, -1, 0], dtype=np.int32)
    tail = b"TT"
    want = b"alpha" + b"j\x01\x02\x03\x04" + tail + b"" + b"h\x07" + tail + b"bc" + b"delta!" + b"h\x07" + tail

    def join(n, ref_of, cap):
        out = bytearray(max(cap, 1) + 8)
        w = lib.pwgs_pickle_join(n, blob, off.ctypes.data, ref_of.ctypes.data_as(C.POINTER(C.c_int32)), 2, refs.ctypes.data, tail, len(tail),
                                 (C.c_char * len(out)).from_buffer(out), cap)
        return w, bytes(out)

    w, out = join(4, ref_of, 64)
    assert w == len(want) and out[:w] == want and out[w:] == bytes(len(out) - w)
    w, out = join(4, ref_of, len(want))
    assert w == len(want)
    w, out = join(4, ref_of, len(want) - 1)
    assert w == -4 and out[len(want) - 1:] == bytes(len(out) - len(want) + 1)    # PWGS_ELIMIT, the bytes beyond cap untouched
    assert join(4, np.array([2, 0, 0, 0], dtype=np.int32), 64)[0] == -1           # reference index out of range
    assert join(0, ref_of, 0)[0] == 0
