"""oracle.py — TEST INFRASTRUCTURE ONLY.

ctypes front-end to the CPU checkers:
  * ``liboracle.so``            our plain-C restatement of the hot path (pwgs_oracle.c);
  * ``_ref/libpwgs_ref.so``     the UNMODIFIED reference mh.cpp/util.cpp behind ref_harness.cpp
                                (built by oracle/Makefile where /root/reference exists; prebuilt elsewhere).
Plus independent readers/writers of the reference's text formats that
feed the compiled reference.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product never does.
Citations are file:line relative to /root/reference.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SEM_CPP, SEM_PY = 0, 1
RNG_MT, RNG_PHILOX = 0, 1

_i = C.POINTER(C.c_int)
_d = C.POINTER(C.c_double)


class _OrcData(C.Structure):
    _fields_ = [("n_ssm", C.c_int), ("n_cnv", C.c_int), ("ntps", C.c_int),
                ("a", _i), ("d", _i), ("mu_r", _d), ("mu_v", _d),
                ("cnv_ptr", _i), ("cnv_idx", _i), ("cnv_c1", _i), ("cnv_c2", _i)]


class _OrcTree(C.Structure):
    _fields_ = [("K", C.c_int), ("parent", _i), ("node_of_datum", _i), ("phi", _d), ("pi", _d)]


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when the reference sources are present)."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")) or \
            os.path.getmtime(os.path.join(HERE, "liboracle.so")) < os.path.getmtime(os.path.join(HERE, "pwgs_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if os.path.exists("/root/reference/mh.cpp"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L.orc_log_bin_coeff.restype = C.c_double
        L.orc_log_bin_coeff.argtypes = [C.c_int, C.c_int]
        L.orc_logsumexp.restype = C.c_double
        L.orc_u01_from_u64.restype = C.c_double
        L.orc_u01_from_u64.argtypes = [C.c_uint64]
        L.orc_multi_param_post.restype = C.c_double
        L.orc_total_llh.restype = C.c_double
        L.orc_datum_ll_py.restype = C.c_double
        L.orc_philox_accept_uniform.restype = C.c_double
        L.orc_philox_accept_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        _lib = L
    return _lib


def _p(arr, typ):
    return arr.ctypes.data_as(typ)


class Data:
    """Immutable read-count data set (arrays as the C ABI wants them)."""

    def __init__(self, a, d, mu_r, mu_v, n_ssm, cnv_ptr=None, cnv_idx=None, cnv_c1=None, cnv_c2=None):
        self.a = np.ascontiguousarray(a, dtype=np.int32)
        self.d = np.ascontiguousarray(d, dtype=np.int32)
        assert self.a.ndim == 2 and self.a.shape == self.d.shape
        self.D, self.T = self.a.shape
        self.n_ssm = int(n_ssm)
        self.n_cnv = self.D - self.n_ssm
        self.mu_r = np.ascontiguousarray(mu_r, dtype=np.float64)
        self.mu_v = np.ascontiguousarray(mu_v, dtype=np.float64)
        if cnv_ptr is None:
            cnv_ptr = np.zeros(self.n_ssm + 1, dtype=np.int32)
            cnv_idx = cnv_c1 = cnv_c2 = np.zeros(0, dtype=np.int32)
        self.cnv_ptr = np.ascontiguousarray(cnv_ptr, dtype=np.int32)
        self.cnv_idx = np.ascontiguousarray(cnv_idx, dtype=np.int32)
        self.cnv_c1 = np.ascontiguousarray(cnv_c1, dtype=np.int32)
        self.cnv_c2 = np.ascontiguousarray(cnv_c2, dtype=np.int32)
        # keep at least one element so the pointers are valid
        self._pad = [np.zeros(1, dtype=np.int32)]

    def cstruct(self):
        def ip(x):
            return _p(x if x.size else self._pad[0], _i)
        return _OrcData(self.n_ssm, self.n_cnv, self.T, ip(self.a), ip(self.d), _p(self.mu_r, _d), _p(self.mu_v, _d),
                        ip(self.cnv_ptr), ip(self.cnv_idx), ip(self.cnv_c1), ip(self.cnv_c2))

    @property
    def cnv_affected(self):
        return np.nonzero(np.diff(self.cnv_ptr) > 0)[0]


class Tree:
    def __init__(self, parent, node_of_datum, phi, pi):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.node_of_datum = np.ascontiguousarray(node_of_datum, dtype=np.int32)
        self.phi = np.ascontiguousarray(phi, dtype=np.float64)
        self.pi = np.ascontiguousarray(pi, dtype=np.float64)
        self.K = len(self.parent)
        assert self.phi.shape == self.pi.shape and self.phi.shape[0] == self.K

    def cstruct(self):
        return _OrcTree(self.K, _p(self.parent, _i), _p(self.node_of_datum, _i), _p(self.phi, _d), _p(self.pi, _d))


# ------------------------------------------------------------------ oracle entry points
def log_bin_coeff(n, k):
    return lib().orc_log_bin_coeff(int(n), int(k))


def lbc_table(data):
    out = np.zeros((data.D, data.T))
    ds = data.cstruct()
    lib().orc_lbc_table(C.byref(ds), _p(out, _d))
    return out


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def data_states(data, tree):
    """params.py:114-171 -> (ssm_of_m [M], coef [M,4,K,2] int32)."""
    ds, ts = data.cstruct(), tree.cstruct()
    M = lib().orc_data_states(C.byref(ds), C.byref(ts), None, None)
    ssm = np.zeros(max(M, 1), dtype=np.int32)
    coef = np.zeros((max(M, 1), 4, tree.K, 2), dtype=np.int32)
    lib().orc_data_states(C.byref(ds), C.byref(ts), _p(ssm, _i), _p(coef, _i))
    return ssm[:M], coef[:M]


def compute_n_genomes(data, tree, j, s, tp):
    ds, ts = data.cstruct(), tree.cstruct()
    nr = np.zeros(4)
    nv = np.zeros(4)
    ns = lib().orc_compute_n_genomes(C.byref(ds), C.byref(ts), int(j), int(s), int(tp), _p(tree.pi, _d), _p(nr, _d), _p(nv, _d))
    return ns, nr, nv


def multi_param_post(data, tree, phi=None, pi=None):
    """mh.cpp:147-166 at (phi, pi) (defaults: the tree's own)."""
    ds, ts = data.cstruct(), tree.cstruct()
    phi = tree.phi if phi is None else np.ascontiguousarray(phi, dtype=np.float64)
    pi = tree.pi if pi is None else np.ascontiguousarray(pi, dtype=np.float64)
    return lib().orc_multi_param_post(C.byref(ds), C.byref(ts), _p(phi, _d), _p(pi, _d))


def llh_rows(data, tree, rows, semantics):
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    out = np.zeros((len(rows), tree.K))
    ds, ts = data.cstruct(), tree.cstruct()
    lib().orc_llh_rows(C.byref(ds), C.byref(ts), len(rows), _p(rows, _i), int(semantics), _p(out, _d))
    return out


def llh_block(data, tree, rows, cols, semantics):
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    out = np.zeros((len(rows), len(cols)))
    ds, ts = data.cstruct(), tree.cstruct()
    lib().orc_llh_block(C.byref(ds), C.byref(ts), len(rows), _p(rows, _i), len(cols), _p(cols, _i), int(semantics), _p(out, _d))
    return out


def total_llh(data, tree, semantics):
    ds, ts = data.cstruct(), tree.cstruct()
    return lib().orc_total_llh(C.byref(ds), C.byref(ts), int(semantics))


def mh_loop(data, tree, iters, mh_std, rng=RNG_PHILOX, seed=0, stream=0, recompute_old=False, want_trace=False):
    """mh.cpp:65-109.  Returns dict(accepted, ratio, phi, pi, llh[, trace])."""
    ds, ts = data.cstruct(), tree.cstruct()
    phi = np.zeros_like(tree.phi)
    pi = np.zeros_like(tree.pi)
    llh = C.c_double(0)
    trace = np.zeros((max(iters, 1), 2)) if want_trace else None
    acc = lib().orc_mh_loop(C.byref(ds), C.byref(ts), int(iters), C.c_double(mh_std), int(rng),
                            C.c_uint64(seed), C.c_uint64(stream), int(recompute_old),
                            _p(phi, _d), _p(pi, _d), C.byref(llh), _p(trace, _d) if want_trace else None)
    out = dict(accepted=acc, ratio=acc / max(iters, 1), phi=phi, pi=pi, llh=llh.value)
    if want_trace:
        out["trace"] = trace[:iters]
    return out


def philox_proposal(tree, mh_std, seed, stream, it):
    pi1 = np.zeros_like(tree.pi)
    phi1 = np.zeros_like(tree.pi)
    T = tree.pi.shape[1]
    lib().orc_philox_proposal(tree.K, T, _p(tree.parent, _i), _p(tree.pi, _d), C.c_double(mh_std),
                              C.c_uint64(seed), C.c_uint64(stream), C.c_uint32(it), _p(pi1, _d), _p(phi1, _d))
    return pi1, phi1


def accept_uniform(seed, stream, it):
    return lib().orc_philox_accept_uniform(seed, stream, it)


# ------------------------------------------------------------------ reference text formats (Appendix B)
def read_ssm_cnv(ssm_fn, cnv_fn=None):
    """util2.py:53-94 / mh.cpp:208-305 -> Data.  SSM ids must be s0,s1,.. in row order (README.md:23-25)."""
    a, d, mu_r, mu_v = [], [], [], []
    with open(ssm_fn) as f:
        hdr = f.readline().rstrip("\n").split("\t")
        for line in f:
            t = line.rstrip("\n").split("\t")
            if len(t) < 4:
                continue
            row = dict(zip(hdr, t))
            assert row["id"] == "s%d" % len(a), "SSM ids must be s0,s1,... in order"
            a.append([int(x) for x in row["a"].split(",")])
            d.append([int(x) for x in row["d"].split(",")])
            mu_r.append(float(row["mu_r"]))
            mu_v.append(float(row["mu_v"]))
    n_ssm = len(a)
    links = [[] for _ in range(n_ssm)]
    n_cnv = 0
    if cnv_fn and os.path.exists(cnv_fn) and os.path.getsize(cnv_fn) > 0:
        with open(cnv_fn) as f:
            hdr = f.readline().rstrip("\n").split("\t")
            for line in f:
                t = line.rstrip("\n").split("\t")
                if len(t) < 3:
                    continue
                row = dict(zip(hdr, t))
                assert row["cnv"] == "c%d" % n_cnv
                a.append([int(x) for x in row["a"].split(",")])
                d.append([int(x) for x in row["d"].split(",")])
                mu_r.append(0.999)            # util2.py:84, mh.cpp:285-286
                mu_v.append(0.5)
                ssms = row.get("ssms") or ""
                for tok in [s for s in ssms.split(";") if s]:
                    sid, c1, c2 = tok.split(",")
                    links[int(sid[1:])].append((n_cnv, int(c1), int(c2)))
                n_cnv += 1
    ptr = np.zeros(n_ssm + 1, dtype=np.int32)
    for j in range(n_ssm):
        ptr[j + 1] = ptr[j] + len(links[j])
    flat = [l for ls in links for l in ls]
    idx = np.array([l[0] for l in flat], dtype=np.int32)
    c1 = np.array([l[1] for l in flat], dtype=np.int32)
    c2 = np.array([l[2] for l in flat], dtype=np.int32)
    return Data(np.array(a), np.array(d), mu_r, mu_v, n_ssm, ptr, idx, c1, c2)


def write_ssm_cnv(data, ssm_fn, cnv_fn):
    """Inverse of read_ssm_cnv (so synthetic inputs can be fed to the compiled reference)."""
    with open(ssm_fn, "w") as f:
        f.write("id\tgene\ta\td\tmu_r\tmu_v\n")
        for j in range(data.n_ssm):
            f.write("s%d\tg%d\t%s\t%s\t%r\t%r\n" % (j, j, ",".join(map(str, data.a[j])), ",".join(map(str, data.d[j])),
                                                  float(data.mu_r[j]), float(data.mu_v[j])))
    per_cnv = [[] for _ in range(data.n_cnv)]
    for j in range(data.n_ssm):
        for l in range(data.cnv_ptr[j], data.cnv_ptr[j + 1]):
            per_cnv[data.cnv_idx[l]].append("s%d,%d,%d" % (j, data.cnv_c1[l], data.cnv_c2[l]))
    with open(cnv_fn, "w") as f:
        if data.n_cnv:
            f.write("cnv\ta\td\tssms\tphysical_cnvs\n")
            for c in range(data.n_cnv):
                j = data.n_ssm + c
                f.write("c%d\t%s\t%s\t%s\t%s\n" % (c, ",".join(map(str, data.a[j])), ",".join(map(str, data.d[j])),
                                                  ";".join(per_cnv[c]), "chrom=1,start=1,end=2,major_cn=1,minor_cn=1,cell_prev=0.5"))


def post_order(parent):
    """Node indices children-before-parent, children in increasing index (params.py:78-99 descend order)."""
    K = len(parent)
    kids = [[] for _ in range(K)]
    root = -1
    for k in range(K):
        if parent[k] < 0:
            root = k
        else:
            kids[parent[k]].append(k)
    out = []

    def descend(n):
        for c in kids[n]:
            descend(c)
        out.append(n)
    descend(root)
    return out


def write_c_tree(tree, fn, order=None, fmt="%.17g"):
    """params.py:68-102.  One line per node in `order` (default post-order).  fmt='%.12g' mimics py2 str(float)."""
    K, parent = tree.K, tree.parent
    order = post_order(parent) if order is None else order
    kids = [[c for c in range(K) if parent[c] == k] for k in range(K)]
    depth = [0] * K
    for k in range(K):
        v, dd = parent[k], 0
        while v >= 0:
            dd += 1
            v = parent[v]
        depth[k] = dd
    with open(fn, "w") as f:
        for k in order:
            dids = [str(j) for j in np.nonzero(tree.node_of_datum == k)[0]]
            f.write("\t".join([str(k), ",".join(fmt % x for x in tree.phi[k]), ",".join(fmt % x for x in tree.pi[k]),
                               str(len(kids[k])), ",".join(map(str, kids[k])) or "-1",
                               str(len(dids)), ",".join(dids) or "-1", str(depth[k])]) + "\n")


def write_c_data_states(ssm_of_m, coef, fn):
    """params.py:167-168: one line per (CNV-affected SSM, node): ssm_idx \\t st1 \\t st2 \\t st3 \\t st4 \\t"""
    M, _, K, _ = coef.shape if len(coef) else (0, 4, 0, 2)
    with open(fn, "w") as f:
        for m in range(M):
            for k in range(K):
                f.write(str(int(ssm_of_m[m])) + "\t" + "\t".join("%d,%d,%d" % (k, coef[m, q, k, 0], coef[m, q, k, 1]) for q in range(4)) + "\t\n")


# ------------------------------------------------------------------ the compiled, unmodified reference
_ref = None


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libpwgs_ref.so"))


def ref_lib():
    global _ref
    if _ref is None:
        # RTLD_DEEPBIND: the reference .so carries its own (static) libstdc++ and GSL-shim symbols; without it they
        # get interposed by same-named symbols already in the Python process and the reference mh_loop crashes.
        R = C.CDLL(os.path.join(HERE, "_ref", "libpwgs_ref.so"), mode=os.RTLD_NOW | os.RTLD_DEEPBIND)
        R.ref_load.restype = C.c_void_p
        R.ref_load.argtypes = [C.c_char_p] * 4 + [C.c_int, C.c_double] + [C.c_int] * 4
        R.ref_free.argtypes = [C.c_void_p]
        R.ref_multi_param_post.restype = C.c_double
        R.ref_multi_param_post.argtypes = [C.c_void_p, C.c_int]
        R.ref_param_post.restype = C.c_double
        R.ref_param_post.argtypes = [C.c_void_p, C.c_int, C.c_int]
        R.ref_datum_ll.restype = C.c_double
        R.ref_datum_ll.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int]
        R.ref_log_bin_coeff.restype = C.c_double
        R.ref_log_bin_coeff.argtypes = [C.c_int, C.c_int]
        R.ref_logsumexp.restype = C.c_double
        R.ref_mh_loop.argtypes = [C.c_void_p, C.c_char_p]
        R.ref_get_params.argtypes = [C.c_void_p, _i, _d, _d]
        _ref = R
    return _ref


class RefState:
    """The reference's own structs loaded through its own text loaders (mh.cpp:208-463)."""

    def __init__(self, data, tree, workdir, ssm_of_m=None, coef=None, order=None, mh_itr=0, mh_std=100.0, fmt="%.17g"):
        self.data, self.tree = data, tree
        self.order = post_order(tree.parent) if order is None else list(order)
        fs = [os.path.join(workdir, n) for n in ("ssm.txt", "cnv.txt", "c_tree.txt", "c_data_states.txt")]
        write_ssm_cnv(data, fs[0], fs[1])
        write_c_tree(tree, fs[2], self.order, fmt)
        if coef is None:
            ssm_of_m, coef = data_states(data, tree)
        write_c_data_states(ssm_of_m, coef, fs[3])
        self.files = fs
        self.workdir = workdir
        self.h = ref_lib().ref_load(*[f.encode() for f in fs], mh_itr, mh_std, data.n_ssm, data.n_cnv, tree.K, data.T)

    def multi_param_post(self, old=1):
        return ref_lib().ref_multi_param_post(self.h, old)

    def param_post(self, tp, old=1):
        return ref_lib().ref_param_post(self.h, old, tp)

    def datum_ll(self, j, phi, tp, old=1):
        return ref_lib().ref_datum_ll(self.h, j, phi, old, tp)

    def mh_loop(self):
        ar = os.path.join(self.workdir, "c_mh_ar.txt")
        ref_lib().ref_mh_loop(self.h, ar.encode())
        ids = np.zeros(self.tree.K, dtype=np.int32)
        phi = np.zeros_like(self.tree.phi)
        pi = np.zeros_like(self.tree.pi)
        ref_lib().ref_get_params(self.h, _p(ids, _i), _p(phi, _d), _p(pi, _d))
        out_phi = np.zeros_like(phi)
        out_pi = np.zeros_like(pi)
        out_phi[ids] = phi
        out_pi[ids] = pi
        return dict(ratio=float(open(ar).read()), phi=out_phi, pi=out_pi)

    def close(self):
        if self.h:
            ref_lib().ref_free(self.h)
            self.h = None
