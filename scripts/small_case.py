"""MH kernel time on small inputs vs CTAs per chain (latency-bound regime)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phylowgs_b200 as P
from oracle import oracle as O
from tests import cases
def run(name, data, tree, std):
    with cases.engine_for(P, data, tree) as eng:
        for ctas in (0, 32, 8, 1):
            eng.set_option("mh_ctas", ctas)
            for col in (0, 1):
                eng.set_option("mh_collapse", col)
                eng.mh_launch(5000, std, 1, 1); eng.mh_wait(fetch=False)
                ms = []
                for k in range(5):
                    eng.mh_launch(5000, std, 1, 10 + k); r = eng.mh_wait(fetch=False); ms.append(r["kernel_ms"])
                print("%-10s ctas=%3d collapse=%d: %.3f ms / 5000 it (acc %.3f)" % (name, ctas, col, np.mean(ms), r["ratio"]), flush=True)
data, tree = cases.bundled(O, True)
run("bundled", data, tree, 100.0)
data, tree = cases.synthetic_case(O, n_ssm=1000, n_cnv=0, T=1, K=9, seed=14)
run("1k", data, tree, 100.0)
data, tree = cases.synthetic_case(O, n_ssm=5000, n_cnv=30, T=1, K=12, seed=15)
run("5k+cnv", data, tree, 400.0)
