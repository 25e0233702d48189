"""-m "not gpu": the Python-3 host glue (phylowgs_b200/host) driven by the CPU oracle backend: it must run the reference's
per-iteration sequence (evolve.py:148-230), keep the TSSB invariants, and write/reload the reference's file formats."""
import os
import pickle
import zipfile

import numpy as np
import pytest

from tests import cases, chain_util
from tests.oracle_backend import OracleBackend

SSM = os.path.join(cases.GOLDEN, "ssm_data.txt")
CNV = os.path.join(cases.GOLDEN, "cnv_data.txt")


def check_invariants(tssb):
    weights, nodes = tssb.get_mixture()
    assert tssb.root["node"].num_local_data() == 0                                   # the root never holds data (tssb.py:118-120)
    assert sum(nd.num_local_data() for nd in nodes) == tssb.num_data
    for n, nd in enumerate(tssb.assignments):
        assert n in nd.data
    assert abs(sum(weights) - 1.0) < 1.0 and all(w >= 0 for w in weights)
    for nd in nodes:                                                                  # params = pi + sum of children's params
        want = np.asarray(nd.pi) + sum((np.asarray(c.params) for c in nd.children()), np.zeros_like(np.asarray(nd.pi)))
        assert np.allclose(nd.params, want, atol=1e-9)
    assert np.allclose(sum(np.asarray(nd.pi) for nd in nodes), 1.0, atol=1e-9)


def test_load_data_bundled_example():
    from phylowgs_b200.host import evolve as E
    data, n_ssms, n_cnvs, mapping = E.load_data(SSM, CNV)
    assert (n_ssms, n_cnvs) == (11, 1) and [d.id for d in data][:3] == ["s0", "s1", "s2"] and data[-1].id == "c0"
    assert [(c.id, a, b) for c, a, b in data[2].cnv] == [("c0", 1, 2)] and [(c.id, a, b) for c, a, b in data[4].cnv] == [("c0", 0, 1)]
    assert data[-1].mu_r == 0.999 and data[-1].mu_v == 0.5 and "c0" in mapping
    assert data[0]._log_bin_norm_const[0] == pytest.approx(87743.31054373167, rel=1e-12)


def test_chain_runs_and_keeps_invariants():
    tssb, n_ssms, n_cnvs = chain_util.new_chain(SSM, CNV, OracleBackend, seed=7)
    rec = chain_util.run_chain(tssb, n_ssms, n_cnvs, iterations=6, mh_iters=200, seed=7)
    check_invariants(tssb)
    assert all(np.isfinite(r["llh"]) for r in rec)
    assert rec[-1]["llh"] > rec[0]["llh"] - 1e6
    # same seed => same chain (host RNG and Philox streams are both seeded)
    tssb2, _, _ = chain_util.new_chain(SSM, CNV, OracleBackend, seed=7)
    rec2 = chain_util.run_chain(tssb2, n_ssms, n_cnvs, iterations=6, mh_iters=200, seed=7)
    for a, b in zip(rec, rec2):
        assert np.array_equal(a["assign"], b["assign"]) and a["llh"] == b["llh"]


def test_metropolis_keeps_reference_contract():
    from phylowgs_b200.host.params import metropolis
    tssb, n_ssms, n_cnvs = chain_util.new_chain(SSM, CNV, OracleBackend, seed=3)
    tssb.resample_assignments()
    tssb.cull_tree()
    before = [np.array(nd.params) for nd in tssb.get_nodes()]
    ar = metropolis(tssb, 300, 100.0, 0, n_ssms, n_cnvs, SSM, CNV, 1, 5, ".")
    assert isinstance(ar, str) and 0.0 <= float(ar) <= 1.0                            # evolve.py:201 does float(state['mh_acc'])
    for nd in tssb.get_nodes():
        assert isinstance(nd.params, np.ndarray) and nd.params.shape == (5,) and nd.pi.shape == (5,)
    assert float(ar) > 0 and any(not np.array_equal(b, nd.params) for b, nd in zip(before, tssb.get_nodes()))
    with pytest.raises(ValueError):
        metropolis(tssb, 10, 100.0, 0, n_ssms + 1, n_cnvs, SSM, CNV, 1, 5, ".")


def test_tree_pickles_use_reference_module_names(tmp_path):
    from phylowgs_b200.host import evolve as E
    E.install_pickle_aliases()
    tssb, n_ssms, n_cnvs = chain_util.new_chain(SSM, CNV, OracleBackend, seed=5)
    chain_util.run_chain(tssb, n_ssms, n_cnvs, iterations=2, mh_iters=50, seed=5)
    blob = pickle.dumps(tssb, protocol=2)
    import pickletools
    globals_ = {arg for op, arg, _ in pickletools.genops(blob) if op.name == "GLOBAL"}
    assert {"tssb TSSB", "alleles alleles", "data Datum"} <= globals_
    assert not any(g.startswith("phylowgs_b200") for g in globals_)
    os.chdir(tmp_path)
    w = E.TreeWriter()
    w.add_extra_file("params.json", "{}")
    w.write_trees([(blob, -1, -12.5), (blob, 0, -11.25)])
    with zipfile.ZipFile("trees.zip") as z:
        assert sorted(z.namelist()) == ["burnin_-1_-12.5", "params.json", "tree_0_-11.25"]
        back = pickle.loads(z.read("tree_0_-11.25"))
    assert back.num_data == 12 and len(back.get_nodes()) == len(tssb.get_nodes()) and back.backend is None
    for attr in ("root", "data", "assignments", "dp_alpha", "dp_gamma", "alpha_decay", "min_depth", "max_depth"):
        assert hasattr(back, attr)
    nd = back.get_nodes()[1]
    for attr in ("data", "_children", "_parent", "tssb", "params", "pi", "params1", "pi1", "path", "ht", "id"):
        assert hasattr(nd, attr), attr


@pytest.mark.parametrize("files", ["bundled", "sim"])
def test_native_sweep_equals_python_sweep(monkeypatch, tmp_path, files):
    """pwgs_resample_assignments (native host code in the library) and sampler.resample_assignments_py consume the same stream
    of uniforms in the same order: the chains they drive must be identical, spawned nodes, sticks and all."""
    if files == "bundled":
        ssm, cnv = SSM, CNV
    else:
        ssm = os.path.join(cases.GOLDEN, "emptysim.4.100.50.ssm.txt")
        cnv = os.path.join(str(tmp_path), "empty.txt")
        open(cnv, "w").close()
    recs, trees = [], []
    for py in ("1", "0"):
        monkeypatch.setenv("PWGS_PY_SAMPLER", py)
        tssb, n_ssms, n_cnvs = chain_util.new_chain(ssm, cnv, OracleBackend, seed=23)
        recs.append(chain_util.run_chain(tssb, n_ssms, n_cnvs, iterations=7, mh_iters=150, seed=23))
        trees.append(tssb)
        check_invariants(tssb)
    for a, b in zip(*recs):
        assert np.array_equal(a["assign"], b["assign"])
        assert a["n_nodes"] == b["n_nodes"] and a["llh"] == b["llh"] and a["hypers"] == b["hypers"]
        assert np.array_equal(a["phi"], b["phi"])
    wa, na = trees[0].get_mixture()
    wb, nb = trees[1].get_mixture()
    assert wa == wb and [sorted(n.data) for n in na] == [sorted(n.data) for n in nb]


def test_async_writer_runs_jobs_in_order_reports_errors_and_drains():
    import threading
    import time
    from phylowgs_b200.host.evolve import AsyncWriter
    w, log = AsyncWriter(), []
    for k in range(5):
        w.submit(lambda k=k: (time.sleep(0.01), log.append(k)))
    w.wait()
    assert log == [0, 1, 2, 3, 4]

    def boom():
        raise OSError("disk full")
    w.submit(boom)
    with pytest.raises(OSError):
        w.wait()
    w.submit(lambda: log.append("after"))                 # the error was delivered once; the writer keeps working
    w.wait()
    assert log[-1] == "after"
    # a signal handler drains: the write in flight finishes, nothing new starts
    started, release = threading.Event(), threading.Event()
    w.submit(lambda: (started.set(), release.wait(5), log.append("in flight")))
    started.wait(5)
    t = threading.Thread(target=w.stop_and_drain)
    t.start()
    time.sleep(0.05)
    assert t.is_alive()                                    # still draining
    release.set()
    t.join(5)
    assert not t.is_alive() and log[-1] == "in flight"
    late = threading.Thread(target=lambda: w.submit(lambda: log.append("late")), daemon=True)
    late.start()
    late.join(0.3)
    assert late.is_alive() and log[-1] == "in flight"      # blocked for good: the process is exiting
