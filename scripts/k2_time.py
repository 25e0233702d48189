"""Times the grid kernel K2 (pwgs_llh_rows over all data) at config 3 for small and large trees."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phylowgs_b200 as P
from phylowgs_b200 import synth
s = synth.config(3)
eng = P.Engine(device=0, **synth.data_kwargs(s))
rng = np.random.default_rng(1)
D = s["a"].shape[0]
for K in (11, 40, 120):
    if K == 11:
        parent, nod, phi, pi = s["parent"], s["node_of_datum"], s["phi"], s["pi"]
    else:
        parent = np.array([-1] + [int(rng.integers(0, k)) for k in range(1, K)], dtype=np.int32)
        pi = rng.dirichlet(np.ones(K), size=1).T.copy(); phi = pi.copy()
        for k in range(K - 1, 0, -1): phi[parent[k]] += phi[k]
        nod = rng.integers(1, K, size=D).astype(np.int32)
    t0 = time.perf_counter(); eng.set_tree(parent, nod, phi, pi); t1 = time.perf_counter()
    buf = eng.host_buffer((K, D))                    # the sweep's form: column-major, into page-locked memory
    eng.llh_range(0, D, None, 1, col_major=True, out=buf)
    ts = []
    for _ in range(3):
        t = time.perf_counter(); g = eng.llh_range(0, D, None, 1, col_major=True, out=buf); ts.append(time.perf_counter() - t)
    sorted_ns = eng.get_option("grid_kernel_ns")
    t = time.perf_counter(); tot = eng.total_llh(1); t2 = time.perf_counter() - t
    eng.set_option("grid_plain", 1)                  # the same grid with the data walked in file order
    eng.llh_range(0, D, None, 1, col_major=True, out=buf)
    plain_ns = eng.get_option("grid_kernel_ns")
    eng.set_option("grid_plain", 0)
    print("K=%3d: file-order kernel %.3f ms" % (K, plain_ns * 1e-6))
    print("K=%3d: set_tree %.2f ms, grid D x K = %d x %d: %.2f ms end to end, kernel %.3f ms (%.1f MB out), total_llh %.2f ms" % (K, 1e3*(t1-t0), D, K, 1e3*min(ts), sorted_ns * 1e-6, g.nbytes/1e6, 1e3*t2))
