"""Config-5 style sweep of the MH kernel: large inputs (slices no longer fit shared memory -> in-place global reads)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phylowgs_b200 as P
from phylowgs_b200 import synth
for (n, T, fc) in ((100000, 1, 0.3), (500000, 1, 0.3), (500000, 5, 0.0), (200000, 5, 0.3)):
    t0 = time.time()
    s = synth.make_synthetic(n, 300 if fc else 0, T, 8, seed=5, frac_cnv=fc, depth=60)
    eng = P.Engine(device=0, **synth.data_kwargs(s))
    eng.set_tree(s["parent"], s["node_of_datum"], s["phi"], s["pi"])
    t1 = time.time()
    for col in (0, 1):
        eng.set_option("mh_collapse", col)
        eng.mh_launch(1000, 100.0, 1, 1); r = eng.mh_wait(fetch=False)
        eng.mh_launch(1000, 100.0, 1, 2); r = eng.mh_wait(fetch=False)
        print("N=%d T=%d cnv_frac=%.1f collapse=%d: setup %.1fs, 1000 it in %.2f ms -> %.0f it/s (acc %.3f, llh %.6g)"
              % (n, T, fc, col, t1 - t0, r["kernel_ms"], 1000 / (r["kernel_ms"] * 1e-3), r["ratio"], r["llh"]), flush=True)
    eng.close()
