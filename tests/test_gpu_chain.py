"""-m gpu: end-to-end chains.  The host glue is deterministic given (host seed, Philox seed), so a chain driven by the CUDA
library must reproduce the chain driven by the CPU oracle: identical assignment indices at every iteration (north_star:
"assignment indices must match bit-exactly given the same uniform draws"), tree-LLH trace and phi within 1e-9 relative."""
import os

import numpy as np
import pytest

from tests import cases, chain_util
from tests.oracle_backend import OracleBackend

pytestmark = pytest.mark.gpu


def gpu_backend(codes):
    from phylowgs_b200.host.backend import GpuBackend
    return GpuBackend(codes, device=0)


def empty_cnv(tmp_path):
    fn = os.path.join(str(tmp_path), "empty_cnv.txt")
    open(fn, "w").close()                                  # README.md:43-44: an empty file when there are no CNVs
    return fn


@pytest.mark.parametrize("name,iterations,mh_iters", [("bundled", 8, 400), ("sim", 5, 300)])
def test_gpu_chain_reproduces_oracle_chain(gpu_lib, tmp_path, name, iterations, mh_iters):
    if name == "bundled":
        ssm, cnv = os.path.join(cases.GOLDEN, "ssm_data.txt"), os.path.join(cases.GOLDEN, "cnv_data.txt")
    else:
        ssm, cnv = os.path.join(cases.GOLDEN, "emptysim.3.20.10.ssm.txt"), empty_cnv(tmp_path)
    recs = []
    for factory in (OracleBackend, gpu_backend):
        tssb, n_ssms, n_cnvs = chain_util.new_chain(ssm, cnv, factory, seed=11)
        recs.append(chain_util.run_chain(tssb, n_ssms, n_cnvs, iterations, mh_iters, seed=11))
        tssb.backend.close()
    for it, (o, g) in enumerate(zip(*recs)):
        assert np.array_equal(o["assign"], g["assign"]), "assignments diverge at iteration %d" % it
        assert o["n_nodes"] == g["n_nodes"]
        assert o["acc"] == g["acc"]
        assert np.max(np.abs(o["phi"] - g["phi"])) < 1e-9
        assert abs(o["llh"] - g["llh"]) <= 1e-9 * abs(o["llh"])


def test_gpu_chain_recovers_simulated_subclones(gpu_lib, tmp_path):
    """phylowgs_sims emptysim.4.100.50: three cancerous populations of 50 SSMs each at depth 100 (truth is implied by the block
    structure).  A short chain on the GPU must put the posterior-mean cellular prevalence of every SSM
    close to its block's 2*VAF and use a handful of major populations (nodes holding >= 5% of the SSMs, the kind of node the
    reference's post-processing keeps).  Stated tolerance: per-SSM RMSE <= 0.06, modal number of major populations in 2..5."""
    ssm, cnv = os.path.join(cases.GOLDEN, "emptysim.4.100.50.ssm.txt"), empty_cnv(tmp_path)
    tssb, n_ssms, n_cnvs = chain_util.new_chain(ssm, cnv, gpu_backend, seed=3)
    rows = [l.split("\t") for l in open(ssm).read().strip().split("\n")[1:]]
    a = np.array([int(r[2]) for r in rows], dtype=float)
    d = np.array([int(r[3]) for r in rows], dtype=float)
    truth = np.repeat([2 * (1 - a[i:i + 50] / d[i:i + 50]).mean() for i in (0, 50, 100)], 50)
    burn, keep = 120, 120
    phi_sum, populated = np.zeros(n_ssms), []
    mh_std = 100.0
    from phylowgs_b200.host.params import metropolis
    for it in range(burn + keep):
        tssb.resample_assignments()
        tssb.cull_tree()
        acc = float(metropolis(tssb, 1000, mh_std, 0, n_ssms, n_cnvs, "", "", 3, 1, ".", stream=it))
        if acc < 0.08 and mh_std < 10000:
            mh_std *= 2.0
        if 0.5 < acc < 0.99:
            mh_std /= 2.0
        tssb.resample_sticks()
        tssb.resample_stick_orders()
        tssb.resample_hypers()
        if it >= burn:
            phi_sum += np.array([float(tssb.assignments[n].params[0]) for n in range(n_ssms)])
            populated.append(sum(1 for nd in tssb.get_nodes() if nd.num_local_data() >= 0.05 * n_ssms))
    tssb.backend.close()
    post = phi_sum / keep
    rmse = float(np.sqrt(np.mean((post - truth) ** 2)))
    mode = int(np.bincount(populated).argmax())
    assert rmse <= 0.06, (rmse, post[::25], truth[::25])
    assert 2 <= mode <= 5, populated


def test_library_side_refreshes_equal_the_python_callbacks(gpu_lib, tmp_path, monkeypatch):
    """The native sweep refreshing the grid itself (pwgs_sweep.ctx: spawned nodes' columns by strided DMA, CNV moves) against
    the same sweep asking Python to do it through its callbacks: same chain, bit for bit, incl. the LLH trace."""
    ssm, cnv = os.path.join(cases.GOLDEN, "ssm_data.txt"), os.path.join(cases.GOLDEN, "cnv_data.txt")
    recs = []
    for callbacks, spare in (("1", None), ("0", None), ("0", "0")):         # spare 0: spawned nodes' columns outside the grid buffer
        monkeypatch.setenv("PWGS_SWEEP_CALLBACKS", callbacks)
        if spare is not None:
            monkeypatch.setenv("PWGS_GRID_SPARE", spare)
        tssb, n_ssms, n_cnvs = chain_util.new_chain(ssm, cnv, gpu_backend, seed=5)
        recs.append(chain_util.run_chain(tssb, n_ssms, n_cnvs, 10, 300, seed=5))
        tssb.backend.close()
    assert sum(r["refreshes"] for r in recs[1]) > 10                        # tree edits did happen
    for it, (a, b) in enumerate(zip(recs[1], recs[2])):
        assert np.array_equal(a["assign"], b["assign"]) and a["llh"] == b["llh"] and a["refreshes"] == b["refreshes"]
    for it, (a, b) in enumerate(zip(recs[0], recs[1])):
        assert np.array_equal(a["assign"], b["assign"]), "assignments diverge at iteration %d" % it
        assert a["refreshes"] == b["refreshes"] and a["acc"] == b["acc"] and a["llh"] == b["llh"]
        assert np.array_equal(a["phi"], b["phi"])


def test_one_tree_sample_at_the_config3_shape(gpu_lib, oracle, tmp_path):
    """BASELINE.json configs[2] (50k SSMs, 300 CNVs) through one whole tree sample from the one-node start: a sweep that spawns a
    hundred nodes and moves dozens of CNVs (every refresh done by the library against the device), 60 MH iterations, the LLH
    trace -- the assignment of every one of the 50 300 data must be the oracle chain's."""
    from phylowgs_b200 import synth
    s = synth.config(3)
    data = oracle.Data(**synth.data_kwargs(s))
    ssm, cnv = str(tmp_path / "ssm.txt"), str(tmp_path / "cnv.txt")
    oracle.write_ssm_cnv(data, ssm, cnv)
    recs = []
    for factory in (OracleBackend, gpu_backend):
        tssb, n_ssms, n_cnvs = chain_util.new_chain(ssm, cnv, factory, seed=3)
        assert (n_ssms, n_cnvs) == (50000, 300)
        recs.append(chain_util.run_chain(tssb, n_ssms, n_cnvs, 1, 60, seed=3))
        tssb.backend.close()
    o, g = recs[0][0], recs[1][0]
    assert o["refreshes"] == g["refreshes"] and o["refreshes"] > 50
    assert np.array_equal(o["assign"], g["assign"]) and o["n_nodes"] == g["n_nodes"] and o["n_nodes"] > 50
    assert o["acc"] == g["acc"]
    assert np.max(np.abs(o["phi"] - g["phi"])) < 1e-9
    assert abs(o["llh"] - g["llh"]) <= 1e-9 * abs(o["llh"])
