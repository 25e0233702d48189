"""Runs a short chain of the host glue (phylowgs_b200/host) on a given backend and records everything a parity test compares."""
import numpy as np

from phylowgs_b200.host import evolve as E
from phylowgs_b200.host.model import TSSB, alleles, boundbeta
from phylowgs_b200.host.params import metropolis


def new_chain(ssm_fn, cnv_fn, backend_factory, seed):
    codes, n_ssms, n_cnvs, _ = E.load_data(ssm_fn, cnv_fn)
    rng = np.random.RandomState(seed)
    backend = backend_factory(codes)
    root = alleles(conc=0.1, ntps=len(codes[0].a))
    tssb = TSSB(dp_alpha=25.0, dp_gamma=1.0, alpha_decay=0.25, root_node=root, data=codes, rng=rng, backend=backend)
    tssb.root["sticks"] = np.vstack([tssb.root["sticks"], .999999])
    child = root.spawn()
    tssb.root["children"].append({"node": child, "main": boundbeta(rng, 1.0, tssb.alpha_decay * tssb.dp_alpha),
                                  "sticks": np.empty((0, 1)), "children": []})
    for n in range(tssb.num_data):
        root.remove_datum(n)
        child.add_datum(n)
        tssb.assignments[n] = child
    for d in codes:
        d.tssb = tssb
    return tssb, n_ssms, n_cnvs


def run_chain(tssb, n_ssms, n_cnvs, iterations, mh_iters, seed, mh_std=100.0):
    """The loop body of evolve.py:148-230.  Returns per-iteration records."""
    out = []
    ntps = len(tssb.data[0].a)
    for it in range(iterations):
        tssb.resample_assignments()
        tssb.cull_tree()
        _, nodes = tssb.get_mixture()
        for i, nd in enumerate(nodes):
            nd.id = i
        E.set_node_height(tssb)
        E.set_path_from_root_to_node(tssb)
        if it == 0:
            E.map_datum_to_node(tssb)
        else:                                              # evolve.do_mcmc relies on the sweeps keeping datum.node in step
            assert all(d.node is tssb.assignments[n] for n, d in enumerate(tssb.data))
        pos = {id(nd): k for k, nd in enumerate(nodes)}
        assign = np.array([pos[id(nd)] for nd in tssb.assignments])
        acc = float(metropolis(tssb, mh_iters, mh_std, 0, n_ssms, n_cnvs, "", "", seed, ntps, ".", stream=it))
        if acc < 0.08 and mh_std < 10000:
            mh_std *= 2.0
        if 0.5 < acc < 0.99:
            mh_std /= 2.0
        phi = np.array([nd.params for nd in nodes])
        tssb.resample_sticks()
        tssb.resample_stick_orders()
        tssb.resample_hypers()
        llh = tssb.complete_data_log_likelihood()
        out.append(dict(assign=assign, n_nodes=len(nodes), acc=acc, phi=phi, llh=llh, refreshes=tssb.last_grid_refreshes,
                        hypers=(tssb.dp_alpha, tssb.dp_gamma, tssb.alpha_decay)))
    return out
