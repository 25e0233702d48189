"""MH kernel time on a wide tree (the trees of burn-in sweeps have 100-200 nodes): node tables in shared memory (small batches)
against node tables in global memory (mh_big = 1, full batches).   python scripts/mh_big_tree.py [K] [collapse]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phylowgs_b200 as P
from phylowgs_b200 import synth
K = int(sys.argv[1]) if len(sys.argv) > 1 else 128
collapse = int(sys.argv[2]) if len(sys.argv) > 2 else 1
s = synth.config(3)
eng = P.Engine(device=0, **synth.data_kwargs(s))
rng = np.random.default_rng(1)
D = s["a"].shape[0]
parent = np.array([-1] + [int(rng.integers(0, k)) for k in range(1, K)], dtype=np.int32)
pi = rng.dirichlet(np.ones(K), size=1).T.copy(); phi = pi.copy()
for k in range(K - 1, 0, -1): phi[parent[k]] += phi[k]
nod = rng.integers(1, K, size=D).astype(np.int32)
eng.set_tree(parent, nod, phi, pi)
eng.set_option("mh_collapse", collapse)
for big in (0, 1):
    for batch in (256,):
        eng.set_option("mh_big", big); eng.set_option("mh_batch", batch)
        ms = []
        for i in range(4):
            eng.mh_launch(5000, 1000.0, 1, i); r = eng.mh_wait(fetch=False); ms.append(r["kernel_ms"])
        print("K=%d collapse=%d big=%d batch<=%d: kernel %.3f ms (accept %.4f)" % (K, collapse, big, batch, min(ms[1:]), r["ratio"]), flush=True)
