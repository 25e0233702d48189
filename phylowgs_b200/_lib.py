"""ctypes binding of libpwgs_b200.so (include/pwgs_b200.h).  No torch, no CPU fallback:
if the CUDA library is missing or no B200 is visible, every compute call raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# PWGS_LIB: a tuning build of the same library (phylowgs_b200/build.py -D... -o...) for A/B timing of kernel variants
LIB_PATH = os.environ.get("PWGS_LIB") or os.path.join(HERE, "libpwgs_b200.so")

SEM_MH, SEM_PY = 0, 1
LAYOUT_COLS = 0x100                      # PWGS_LAYOUT_COLS: column-major output of pwgs_llh_rows / pwgs_llh_block

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)

# name -> (restype, argtypes); kept in one table so tests can check it against include/pwgs_b200.h
SIGNATURES = {
    "pwgs_last_error": (C.c_char_p, []),
    "pwgs_version": (C.c_int, []),
    "pwgs_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pwgs_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p,
                              _i32p, _i32p, _i32p, _i32p, C.POINTER(C.c_void_p)]),
    "pwgs_destroy": (C.c_int, [C.c_void_p]),
    "pwgs_set_tree": (C.c_int, [C.c_void_p, C.c_int, _i32p, _i32p, _f64p, _f64p]),
    "pwgs_mh_run": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_uint64, _f64p, _f64p, _f64p, _f64p]),
    "pwgs_mh_launch": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_uint64]),
    "pwgs_mh_launch_multi": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_int, _f64p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pwgs_mh_wait": (C.c_int, [C.c_void_p, _f64p, _f64p, _f64p, _f64p, C.POINTER(C.c_float)]),
    "pwgs_llh_rows": (C.c_int, [C.c_void_p, C.c_int, _i32p, C.c_int, _f64p]),
    "pwgs_llh_block": (C.c_int, [C.c_void_p, C.c_int, _i32p, C.c_int, _i32p, C.c_int, _f64p]),
    "pwgs_llh_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _f64p]),
    "pwgs_llh_range_pitched": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _f64p, C.c_int64]),
    "pwgs_move_data": (C.c_int, [C.c_void_p, C.c_int, _i32p, _i32p]),
    "pwgs_host_buffer": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "pwgs_total_llh": (C.c_int, [C.c_void_p, C.c_int, _f64p]),
    "pwgs_get_lbc": (C.c_int, [C.c_void_p, _f64p]),
    "pwgs_get_data_states": (C.c_int, [C.c_void_p, _i32p, _i32p, _i32p]),
    "pwgs_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "pwgs_get_option": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]),
    "pwgs_get_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "pwgs_resample_assignments": (C.c_int, [C.c_void_p]),      # struct pwgs_sweep*, see host/sampler.py::_Sweep
    "pwgs_pickle_join": (C.c_int64, [C.c_int32, C.c_void_p, C.c_void_p, _i32p, C.c_int32, C.c_void_p, C.c_char_p, C.c_int32,
                                     C.c_void_p, C.c_int64]),
    "pwgs_math_selftest": (C.c_int, [C.c_int, C.c_int, _f64p, _f64p]),
    "pwgs_fp64_peak": (C.c_int, [C.c_int, _f64p]),
    "pwgs_selftest_llh_host": (C.c_int, [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, _i32p, _i32p, _i32p, _i32p,
                                         C.c_int, _i32p, _i32p, _f64p, _f64p, C.c_int, _i32p, C.c_int, _f64p]),
    "pwgs_plan_partition": (C.c_int, [C.c_int, C.c_int, _i32p, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p]),
}

_lib = None


class PwgsError(RuntimeError):
    pass


def load():
    """Load the CUDA library, failing loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PwgsError("%s is missing: build it with `python -m phylowgs_b200.build` "
                            "(there is no CPU fallback for the MH / likelihood path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise PwgsError("libpwgs_b200 error %d: %s" % (rc, load().pwgs_last_error().decode()))


def i32(x):
    return np.ascontiguousarray(x, dtype=np.int32)


def f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def p_i32(x):
    return x.ctypes.data_as(_i32p) if x is not None and x.size else None


def p_f64(x):
    return x.ctypes.data_as(_f64p) if x is not None and x.size else None


def device_count():
    n = C.c_int(0)
    rc = load().pwgs_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def fp64_peak(device=0):
    v = C.c_double(0)
    check(load().pwgs_fp64_peak(device, C.byref(v)))
    return v.value
