"""`params.metropolis` with the reference's signature and return value (params.py:30-65), minus the files and the subprocess:
the current tree goes to the device as arrays (pwgs_set_tree), one persistent kernel runs the `iters` Metropolis-Hastings
iterations (pwgs_mh_run), and the final node.params / node.pi come back in full fp64 (update_tree_params, params.py:183-194)."""
import numpy as np


def get_c_fnames(tmp_dir):
    """Kept for callers that clean up the exchange directory (evolve.py:394-406); nothing is written there any more."""
    import os
    return tuple(os.path.join(tmp_dir, "c_%s.txt" % n) for n in ("tree", "data_states", "params", "mh_ar"))


def metropolis(tssb, iters=1000, std=0.01, burnin=0, n_ssms=0, n_cnvs=0, fin1='', fin2='', rseed=1, ntps=5, tmp_dir='.',
               stream=0):
    """Returns the acceptance ratio as a string, like the reference (the caller does float(state['mh_acc']), evolve.py:201).
    `burnin`, `fin1`, `fin2`, `tmp_dir` are accepted and ignored (the reference ignores burnin too; the others only served the
    file round trip).  `rseed` keys the device's Philox stream together with `stream` (the MCMC iteration), where the
    reference re-seeds MT19937 identically on every call (mh.cpp:66)."""
    backend = tssb.backend
    if backend is None:
        raise RuntimeError("tssb.backend is not set: metropolis() needs the CUDA context of the chain")
    if n_ssms + n_cnvs != tssb.num_data:
        raise ValueError("n_ssms + n_cnvs does not match the number of data")
    nodes = backend.sync_tree(tssb)
    res = backend.mh(int(iters), float(std), int(rseed), int(stream))
    for k, nd in enumerate(nodes):
        nd.params = np.array(res["phi"][k])
        nd.pi = np.array(res["pi"][k])
    tssb.last_mh_llh = res["llh"]
    return repr(float(res["ratio"]))
