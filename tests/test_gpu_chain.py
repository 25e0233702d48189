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
