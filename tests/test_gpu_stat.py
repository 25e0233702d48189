"""-m gpu: end-to-end STATISTICAL parity on a phylowgs_sims file (SURVEY.md 8c last row; north_star: "end-to-end chains on
phylowgs_sims must reproduce the reference's posterior phi means, cluster counts and tree-LLH traces within a stated statistical
tolerance").  Four chains of this repo's evolve driver on cuda:0 (emptysim.4.100.50: 150 SSMs, 100 burn-in + 200 samples, 5000 MH
iterations each) against four chains of the REFERENCE -- its own evolve.py / tssb.py / data.py / params.py executed under Python 3
(oracle/ref_py3.py) driving the unmodified mh.cpp through its own text files, frozen in tests/golden/stat_ref_sim150.json by
scripts/stat_parity.py.  Tolerances (scripts/stat_parity.py::compare_summaries): post-burn-in mean LLH within 3 between-seed
standard deviations, equal modal population count, per-SSM posterior-mean phi RMSE <= 0.02, MH acceptance ratio within 0.05.
The chains are deterministic given their seeds, so this is a regression test, not a coin toss."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "scripts", "stat_parity.py")


def test_four_chains_against_the_frozen_reference_chains(gpu_lib, tmp_path):
    spec = importlib.util.spec_from_file_location("stat_parity", SCRIPT)
    sp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sp)
    dirs, procs = [], []
    for seed in (1, 2, 3, 4):
        d = str(tmp_path / ("s%d" % seed))
        dirs.append(d)
        procs.append(subprocess.Popen([sys.executable, SCRIPT, "gpu", "--set", "sim150", "--seed", str(seed), "-O", d], cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out.decode()[-2000:]
    gpu = dict(chains=[sp.summarise_chain(d) for d in dirs])
    import numpy as np
    gpu["mean_phi"] = np.mean([np.array(c["mean_phi"]) for c in gpu["chains"]], axis=0).tolist()
    with open(os.path.join(ROOT, "tests", "golden", "stat_ref_sim150.json")) as f:
        ref = json.load(f)
    rows, ok = sp.compare_summaries(ref, gpu)
    table = "\n".join("%s | %s | %s | %s | %s" % (q, a, b, tol, "yes" if good else "NO") for q, a, b, tol, good in rows)
    assert ok, table
    assert all(c["n_post"] == 200 and c["n_trees"] == 200 for c in gpu["chains"])
