"""The shared / global-memory switch of the MH kernel's node tables on multi-sample data (config 4: 20k SSMs x 5 samples): kernel
time with mh_big 0 / 1 for trees around K*T = 56.   python scripts/mh_big_tree_t5.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phylowgs_b200 as P
from phylowgs_b200 import synth
s = synth.config(4)
eng = P.Engine(device=0, **synth.data_kwargs(s))
D, T = s["a"].shape
for K in [int(k) for k in (sys.argv[1:] or (9, 11, 12, 14, 18, 26))]:
    rng = np.random.default_rng(K)
    parent = np.array([-1] + [int(rng.integers(0, k)) for k in range(1, K)], dtype=np.int32)
    pi = rng.dirichlet(np.ones(K), size=T).T.copy(); phi = pi.copy()
    for k in range(K - 1, 0, -1): phi[parent[k]] += phi[k]
    nod = rng.integers(1, K, size=D).astype(np.int32)
    eng.set_tree(parent, nod, phi, pi)
    for collapse in (1, 0):
        eng.set_option("mh_collapse", collapse)
        out = []
        for big in (0, 1, -1):
            eng.set_option("mh_big", big)
            ms = []
            for i in range(4):
                eng.mh_launch(5000, 1000.0, 1, i); r = eng.mh_wait(fetch=False); ms.append(r["kernel_ms"])
            out.append("big=%2d %.3f ms" % (big, min(ms[1:])))
        print("K=%2d K*T=%3d collapse=%d: %s (accept %.4f)" % (K, K * T, collapse, "  ".join(out), r["ratio"]), flush=True)
