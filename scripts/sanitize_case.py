"""Small end-to-end use of every kernel for compute-sanitizer (memcheck / racecheck): bundled example with its CNV."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phylowgs_b200 as P
from oracle import oracle as O
from tests import cases
data, tree = cases.synthetic_case(O, n_ssm=400, n_cnv=10, T=2, K=7, seed=21)
with cases.engine_for(P, data, tree) as eng:
    for dist, stage, ctas in ((1, 1, 0), (0, 0, 3)):
        eng.set_option("mh_dist", dist); eng.set_option("mh_stage", stage); eng.set_option("mh_ctas", ctas)
        r = eng.mh_run(300, 1000.0, 1234, 7)
        print("mh", dist, stage, ctas, r["ratio"], r["llh"])
    print(eng.total_llh(P.SEM_PY), eng.llh_rows(np.arange(20), P.SEM_PY).sum(), eng.llh_block(np.arange(5), [1, 2]).sum())
    eng.data_states(); eng.lbc()
print("done")
