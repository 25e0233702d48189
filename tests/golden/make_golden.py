"""Generates tests/golden/ref_vectors.json from the UNMODIFIED reference (oracle/_ref/libpwgs_ref.so = /root/reference/mh.cpp +
util.cpp compiled where they lie, behind oracle/ref_harness.cpp).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Every number in the JSON comes out of the reference's own functions (load_* , param_post, multi_param_post, datum::log_ll,
log_bin_coeff, logsumexp, mh_loop).  Inputs are the seeded cases of tests/cases.py; their SHA-256 is stored so a drift of the
generator is reported as such.  The stochastic vectors (mh_loop) use the GSL stand-in oracle/gsl_shim (libgsl is not available
here): they pin the LOOP LOGIC of mh.cpp:65-142 draw for draw, not GSL's random stream ("parity unpinned" at the GSL boundary).
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O   # noqa: E402
from tests import cases          # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_vectors.json")


def digest(data, tree):
    h = hashlib.sha256()
    for arr in (data.a, data.d, data.mu_r, data.mu_v, data.cnv_ptr, data.cnv_idx, data.cnv_c1, data.cnv_c2,
                tree.parent, tree.node_of_datum, tree.phi, tree.pi):
        h.update(np.ascontiguousarray(arr).tobytes())
    return h.hexdigest()


def vectors_for(name, data, tree, mh_iters=0, mh_std=100.0):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else len(name) * 7919)
    out = {"name": name, "sha256": digest(data, tree), "D": data.D, "T": data.T, "K": tree.K}
    with tempfile.TemporaryDirectory() as td:
        ref = O.RefState(data, tree, td)
        out["multi_param_post"] = ref.multi_param_post()
        out["param_post"] = [ref.param_post(tp) for tp in range(data.T)]
        picks = sorted(set(rng.integers(0, data.D, size=12).tolist() + data.cnv_affected[:12].tolist() + list(range(data.n_ssm, min(data.D, data.n_ssm + 3)))))
        dl = []
        for j in picks:
            for tp in sorted(set([0, data.T - 1])):
                phi = float(tree.phi[tree.node_of_datum[j], tp])
                dl.append([int(j), int(tp), phi, ref.datum_ll(j, phi, tp)])
        out["datum_ll"] = dl
        ref.close()
    if mh_iters:
        # relabel so that index order is a children-first order: the reference then draws its gammas in index order
        t2 = cases.children_first(tree)
        with tempfile.TemporaryDirectory() as td:
            ref = O.RefState(data, t2, td, order=list(range(t2.K)), mh_itr=mh_iters, mh_std=mh_std)
            res = ref.mh_loop()
            ref.close()
        out["mh"] = {"iters": mh_iters, "mh_std": mh_std, "ratio": res["ratio"], "phi": res["phi"].tolist(), "pi": res["pi"].tolist()}
    return out


def main():
    assert O.ref_available(), "oracle/_ref/libpwgs_ref.so missing: make -C oracle ref"
    vec = {"generator": "tests/golden/make_golden.py", "source": "unmodified /root/reference/mh.cpp + util.cpp (g++ -O3) + oracle/gsl_shim",
           "cases": [], "log_bin_coeff": [], "logsumexp": []}
    R = O.ref_lib()
    for n, k in [(126755, 66023), (10, 0), (10, 10), (1, 0), (0, 0), (185970, 1234), (60, 31), (3000000, 1500000)]:
        vec["log_bin_coeff"].append([n, k, R.ref_log_bin_coeff(n, k)])
    import ctypes as C
    for xs in ([0.0, 0.0, 0.0, 0.0], [-1000.0, -1001.0, -999.5, -1200.0], [-227.955924206411, -5.0, -227.955924206411, -5.5], [3.0]):
        arr = (C.c_double * len(xs))(*xs)
        vec["logsumexp"].append([xs, R.ref_logsumexp(arr, len(xs))])
    for with_cnv in (False, True):
        data, tree = cases.bundled(O, with_cnv)
        vec["cases"].append(vectors_for("bundled_cnv" if with_cnv else "bundled_plain", data, tree, mh_iters=400))
    for i, kw in enumerate(cases.GOLDEN_SYNTH):
        data, tree = cases.synthetic_case(O, **kw)
        vec["cases"].append(vectors_for("synth_%d" % i, data, tree, mh_iters=250 if i < 3 else 0, mh_std=[100.0, 1000.0, 30.0][i % 3]))
    with open(OUT, "w") as f:
        json.dump(vec, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
