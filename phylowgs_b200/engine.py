"""Array-level Python face of the C ABI: one `Engine` = one chain's device context.

This is what the reference-shaped host glue (params.metropolis, TSSB.resample_assignments,
TSSB.complete_data_log_likelihood) calls; it carries no algorithm of its own."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import LAYOUT_COLS, SEM_MH, SEM_PY, PwgsError, check, f64, i32, p_f64, p_i32  # noqa: F401


class Engine:
    """Device-resident read-count data of one chain (replaces re-parsing ssm_data.txt / cnv_data.txt in every
    mh.o call, mh.cpp:44-46)."""

    def __init__(self, a, d, mu_r, mu_v, n_ssm, cnv_ptr=None, cnv_idx=None, cnv_c1=None, cnv_c2=None, device=0):
        L = _lib.load()
        self.a, self.d = i32(a), i32(d)
        if self.a.ndim != 2 or self.a.shape != self.d.shape:
            raise ValueError("a and d must be [D, T] arrays of equal shape")
        self.D, self.T = self.a.shape
        self.n_ssm = int(n_ssm)
        self.n_cnv = self.D - self.n_ssm
        self.mu_r, self.mu_v = f64(mu_r), f64(mu_v)
        if self.mu_r.shape != (self.D,) or self.mu_v.shape != (self.D,):
            raise ValueError("mu_r and mu_v must have one entry per datum")
        self.K = 0
        self._h = C.c_void_p()
        ptr = i32(cnv_ptr) if cnv_ptr is not None else None
        idx = i32(cnv_idx) if cnv_idx is not None else None
        c1 = i32(cnv_c1) if cnv_c1 is not None else None
        c2 = i32(cnv_c2) if cnv_c2 is not None else None
        check(L.pwgs_create(device, self.n_ssm, self.n_cnv, self.T, p_i32(self.a), p_i32(self.d), p_f64(self.mu_r),
                            p_f64(self.mu_v), p_i32(ptr), p_i32(idx), p_i32(c1), p_i32(c2), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().pwgs_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ tree
    def set_tree(self, parent, node_of_datum, phi, pi):
        """node_of_datum None: the device keeps the assignments it has (a tree that only grew, see pwgs_move_data)."""
        parent = i32(parent)
        nod = i32(node_of_datum) if node_of_datum is not None else None
        phi, pi = f64(phi), f64(pi)
        K = len(parent)
        if phi.shape != (K, self.T) or pi.shape != (K, self.T):
            raise ValueError("phi and pi must be [K, T]")
        if nod is not None and nod.shape != (self.D,):
            raise ValueError("node_of_datum must have one entry per datum")
        check(_lib.load().pwgs_set_tree(self._h, K, p_i32(parent), p_i32(nod), p_f64(phi), p_f64(pi)))
        self.K = K

    def move_data(self, datum, node):
        datum, node = i32(np.atleast_1d(datum)), i32(np.atleast_1d(node))
        check(_lib.load().pwgs_move_data(self._h, len(datum), p_i32(datum), p_i32(node)))

    def host_buffer(self, shape):
        """float64 array of `shape` in the context's page-locked buffer (valid until the next, larger request)."""
        n = int(np.prod(shape))
        ptr = C.c_void_p()
        check(_lib.load().pwgs_host_buffer(self._h, max(8 * n, 8), C.byref(ptr)))
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(max(n, 1),))[:n].reshape(shape)

    # ------------------------------------------------------------------ MH
    def mh_run(self, iters, mh_std, seed=0, stream=0):
        """-> dict(phi [K,T], pi [K,T], ratio, llh)"""
        phi = np.zeros((self.K, self.T))
        pi = np.zeros((self.K, self.T))
        ratio, llh = C.c_double(0), C.c_double(0)
        check(_lib.load().pwgs_mh_run(self._h, int(iters), float(mh_std), int(seed), int(stream),
                                      p_f64(phi), p_f64(pi), C.byref(ratio), C.byref(llh)))
        return dict(phi=phi, pi=pi, ratio=ratio.value, llh=llh.value)

    def mh_launch(self, iters, mh_std, seed=0, stream=0):
        check(_lib.load().pwgs_mh_launch(self._h, int(iters), float(mh_std), int(seed), int(stream)))

    def mh_wait(self, fetch=True):
        phi = np.zeros((self.K, self.T)) if fetch else None
        pi = np.zeros((self.K, self.T)) if fetch else None
        ratio, llh, ms = C.c_double(0), C.c_double(0), C.c_float(0)
        check(_lib.load().pwgs_mh_wait(self._h, p_f64(phi), p_f64(pi), C.byref(ratio), C.byref(llh), C.byref(ms)))
        return dict(phi=phi, pi=pi, ratio=ratio.value, llh=llh.value, kernel_ms=ms.value)

    @staticmethod
    def mh_launch_multi(engines, iters, mh_std, seeds, streams):
        """Several chains of one device in ONE cooperative launch (pwgs_mh_launch_multi); every engine then takes its own mh_wait()."""
        n = len(engines)
        hs = (C.c_void_p * n)(*[e._h for e in engines])
        std = np.ascontiguousarray(np.broadcast_to(np.asarray(mh_std, dtype=np.float64), (n,)))
        sd = (C.c_uint64 * n)(*[int(x) for x in np.broadcast_to(np.asarray(seeds, dtype=np.uint64), (n,))])
        st = (C.c_uint64 * n)(*[int(x) for x in np.broadcast_to(np.asarray(streams, dtype=np.uint64), (n,))])
        check(_lib.load().pwgs_mh_launch_multi(n, hs, int(iters), p_f64(std), sd, st))

    # ------------------------------------------------------------------ likelihoods
    def llh_rows(self, rows, semantics=SEM_PY, col_major=False):
        """-> [n_rows, K]; col_major: [K, n_rows] (one contiguous column per node, PWGS_LAYOUT_COLS)."""
        rows = i32(rows)
        out = np.empty((self.K, len(rows)) if col_major else (len(rows), self.K))
        check(_lib.load().pwgs_llh_rows(self._h, len(rows), p_i32(rows), int(semantics) | (LAYOUT_COLS if col_major else 0), p_f64(out)))
        return out

    def llh_block(self, rows, cols, semantics=SEM_PY, col_major=False):
        rows, cols = i32(rows), i32(cols)
        out = np.empty((len(cols), len(rows)) if col_major else (len(rows), len(cols)))
        if out.size == 0:
            return out
        check(_lib.load().pwgs_llh_block(self._h, len(rows), p_i32(rows), len(cols), p_i32(cols),
                                         int(semantics) | (LAYOUT_COLS if col_major else 0), p_f64(out)))
        return out

    def llh_range(self, row0, n_rows, cols=None, semantics=SEM_PY, col_major=False, out=None):
        """llh_block for the data row0 .. row0 + n_rows - 1 (cols None: all nodes); `out`: where to write (e.g. host_buffer)."""
        cols = i32(cols) if cols is not None else None
        n_cols = len(cols) if cols is not None else self.K
        shape = (n_cols, n_rows) if col_major else (n_rows, n_cols)
        if out is None:
            out = np.empty(shape)
        assert out.shape == shape and out.flags.c_contiguous and out.dtype == np.float64
        if out.size:
            check(_lib.load().pwgs_llh_range(self._h, int(row0), int(n_rows), n_cols, p_i32(cols),
                                             int(semantics) | (LAYOUT_COLS if col_major else 0), p_f64(out)))
        return out

    def total_llh(self, semantics=SEM_PY):
        v = C.c_double(0)
        check(_lib.load().pwgs_total_llh(self._h, int(semantics), C.byref(v)))
        return v.value

    # ------------------------------------------------------------------ diagnostics
    def lbc(self):
        out = np.zeros((self.D, self.T))
        check(_lib.load().pwgs_get_lbc(self._h, p_f64(out)))
        return out

    def data_states(self):
        M = C.c_int32(0)
        check(_lib.load().pwgs_get_data_states(self._h, C.byref(M), None, None))
        ssm = np.zeros(max(M.value, 1), dtype=np.int32)
        coef = np.zeros((max(M.value, 1), 4, self.K, 2), dtype=np.int32)
        check(_lib.load().pwgs_get_data_states(self._h, C.byref(M), p_i32(ssm), p_i32(coef)))
        return ssm[:M.value], coef[:M.value]

    def stream_handle(self):
        """cudaStream_t of this context as an int (for torch.cuda.ExternalStream in bench.py)."""
        h = C.c_void_p()
        check(_lib.load().pwgs_get_stream(self._h, C.byref(h)))
        return h.value or 0

    def set_option(self, name, value):
        check(_lib.load().pwgs_set_option(self._h, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_int64(0)
        check(_lib.load().pwgs_get_option(self._h, name.encode(), C.byref(v)))
        return v.value
