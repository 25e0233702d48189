"""-m "not gpu": oracle/py_half.py -- the plain-Python restatement of the reference's Python half that bench.py times as the CPU
baseline of a tree sample -- against the reference's OWN Python code through tests/golden/py_vectors.json (produced by executing
data.py / tssb.py under Python 3, oracle/ref_py3.py).  Same likelihoods, same sweep: what is timed is what the reference does."""
import numpy as np
import pytest

from oracle import py_half
from tests import refgold
from tests.test_py_golden import CASES, LLH_RTOL, SWEEPS, check_sweep, digest, host_tree, relerr


@pytest.mark.parametrize("name", sorted(CASES))
def test_restated_logprob_matches_datum_log_likelihood(oracle, name):
    g = CASES[name]
    tssb, nodes, arr = host_tree(oracle, name, g["sticks_seed"])
    want = np.array(g["grid"])
    rows = g["rows"]
    step = max(1, len(rows) // 25)                     # every 25th row of the fixture: the per-call cost is the reference's
    for r in range(0, len(rows), step):
        n = rows[r]
        home = tssb.assignments[n]
        got = []
        for k, nd in enumerate(nodes):                 # the datum sits in the candidate node while it is evaluated (tssb.py:122-124)
            tssb.assignments[n].remove_datum(n)
            nd.add_datum(n)
            tssb.assignments[n] = nd
            got.append(py_half.logprob(tssb, tssb.data[n], nd))
        tssb.assignments[n].remove_datum(n)
        home.add_datum(n)
        tssb.assignments[n] = home
        assert relerr(got, want[r]) < LLH_RTOL, (n, got, want[r])


@pytest.mark.parametrize("name", sorted(SWEEPS))
def test_restated_sweep_reproduces_the_reference_sweep(oracle, name):
    g = SWEEPS[name]
    tssb, _, arr = host_tree(oracle, name, g["sticks_seed"])
    assert digest(arr) == g["sha256"]
    # one array of uniforms, consumed in the reference's order (find_node's spawns take theirs from tssb.ustream by inversion,
    # exactly as make_py_golden.py arranged for the reference's own run)
    tssb.rng = tssb.ustream = refgold.ArrayStream(np.random.RandomState(g["uniform_seed"]).rand(g["n_uniforms"]))
    py_half.sweep_sample(tssb, range(tssb.num_data), tssb.rng)
    tssb.ustream = None
    check_sweep(tssb, g)


def test_restated_llh_trace_matches_complete_data_log_likelihood(oracle):
    name = sorted(CASES)[0]
    g = CASES[name]
    tssb, _, _ = host_tree(oracle, name, g["sticks_seed"])
    _, data_term = py_half.llh_trace_sample(tssb, range(tssb.num_data))
    weights, nodes = tssb.get_mixture()
    total = data_term + sum(nd.num_local_data() * np.log(w) for w, nd in zip(weights, nodes) if nd.num_local_data())
    assert relerr(total, g["cdll"]) < LLH_RTOL
