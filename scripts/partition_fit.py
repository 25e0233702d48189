"""Calibrate the work partition's round weights: run K1 with the phase counters on, read every CTA's likelihood cycles and its
rounds per class (plain thirds, CNV rounds with 1/2/3 merged states), fit cycles ~ sum_class rounds * weight by least squares.
    python scripts/partition_fit.py [config3|config4] [plain,cnv1,cnv2,cnv3]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import phylowgs_b200 as P  # noqa: E402
from phylowgs_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config3"
s, _ = bench.workload(name)
eng = P.Engine(device=0, **synth.data_kwargs(s))
eng.set_tree(s["parent"], s["node_of_datum"], s["phi"], s["pi"])
if len(sys.argv) > 2 and sys.argv[2] != "-v":
    for n, v in zip(("mh_w_plain", "mh_w_cnv1", "mh_w_cnv2", "mh_w_cnv3"), sys.argv[2].split(",")):
        eng.set_option(n, int(v))
eng.set_option("mh_profile", 1)
for k in range(4):
    eng.mh_launch(5000, 100.0, 1, k)
    r = eng.mh_wait(fetch=False)
ncta = eng.get_option("n_sms")
cyc = np.array([eng.get_option("mh_lik_cta_%d" % r) for r in range(ncta)], dtype=float)
X = np.array([[eng.get_option("mh_plan_%d_%d" % (c, r)) for c in range(4)] for r in range(ncta)], dtype=float)
X[:, 0] /= max(1, eng.get_option("mh_plan_classes"))   # fractions -> rounds
use = X.sum(axis=0) > 0
w, *_ = np.linalg.lstsq(X[:, use], cyc, rcond=None)
fit = X[:, use] @ w
print(name, "kernel ms %.4f" % r["kernel_ms"], "cycles: min %.0f mean %.0f max %.0f (max/mean %.3f)" % (cyc.min(), cyc.mean(), cyc.max(), cyc.max() / cyc.mean()))
print("rounds per CTA (mean): plain %.2f cnv1 %.2f cnv2 %.2f cnv3 %.2f" % tuple(X.mean(axis=0)))
names = [n for n, u in zip(("plain", "cnv1", "cnv2", "cnv3"), use) if u]
print("fitted cycles per round:", dict(zip(names, np.round(w, 0))), " relative to plain=100:", dict(zip(names, np.round(100 * w / w[0], 1))))
print("residual rms %.0f cycles (%.2f%% of mean)" % (np.sqrt(np.mean((fit - cyc) ** 2)), 100 * np.sqrt(np.mean((fit - cyc) ** 2)) / cyc.mean()))
if "-v" in sys.argv:
    for k in sorted(set(map(tuple, X.tolist()))):
        sel = np.all(X == np.array(k), axis=1)
        print("rounds (plain, cnv1, cnv2, cnv3) = (%.2f, %d, %d, %d): %3d CTAs, cycles min %.0f mean %.0f max %.0f" % (k + (sel.sum(), cyc[sel].min(), cyc[sel].mean(), cyc[sel].max())))
    order = np.argsort(-cyc)[:4]
    for r_ in order:
        print("slowest CTAs: rank %d cycles %.0f rounds %s" % (r_, cyc[r_], X[r_].tolist()))
