"""The grid likelihood the K2 kernel runs (datum_ll_at with the sparse subtree-delta sums), compiled for the HOST and called
through the diagnostic entry pwgs_selftest_llh_host, against the oracle -- so that the device code's logic is checked where
there is no GPU.  Tolerance 1e-9 relative (north_star); exact-zero cases (states with no variant copies) must agree in kind."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from phylowgs_b200 import _lib
from tests import cases


def host_rows(data, tree, rows, sem):
    L = _lib.load()
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    out = np.zeros((len(rows), tree.K))
    i32, f64 = _lib.p_i32, _lib.p_f64
    a, d = np.ascontiguousarray(data.a, dtype=np.int32), np.ascontiguousarray(data.d, dtype=np.int32)
    _lib.check(L.pwgs_selftest_llh_host(data.n_ssm, data.D - data.n_ssm, data.T, i32(a), i32(d), f64(data.mu_r), f64(data.mu_v),
                                        i32(data.cnv_ptr), i32(data.cnv_idx), i32(data.cnv_c1), i32(data.cnv_c2),
                                        tree.K, i32(tree.parent), i32(tree.node_of_datum), f64(tree.phi), f64(tree.pi),
                                        len(rows), i32(rows), int(sem), f64(out)))
    return out


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


@pytest.mark.parametrize("case", range(len(cases.GOLDEN_SYNTH)))
@pytest.mark.parametrize("sem", [O.SEM_CPP, O.SEM_PY])
def test_host_compiled_grid_likelihood_matches_the_oracle(case, sem):
    data, tree = cases.synthetic_case(O, **cases.GOLDEN_SYNTH[case])
    rows = np.arange(data.D, dtype=np.int32)
    want = O.llh_rows(data, tree, rows, sem)
    got = host_rows(data, tree, rows, sem)
    assert rel(got, want) < 1e-9


@pytest.mark.parametrize("with_cnv", [True, False])
@pytest.mark.parametrize("sem", [O.SEM_CPP, O.SEM_PY])
def test_bundled_example(with_cnv, sem):
    data, tree = cases.bundled(O, with_cnv)
    rows = np.arange(data.D, dtype=np.int32)
    assert rel(host_rows(data, tree, rows, sem), O.llh_rows(data, tree, rows, sem)) < 1e-9


def test_deep_trees_many_links_and_zero_mass_nodes():
    """Chains of nodes, several CNV links per SSM (multi_cnv), nodes with pi = 0: the subtree-delta sums must give exact
    zeros where the dense walk does (the nv == 0 state tests of mh.hpp / data.py branch on them)."""
    for seed in range(6):
        data, tree = cases.synthetic_case(O, n_ssm=120, n_cnv=14, T=2, K=5 + 7 * seed, seed=50 + seed, frac_cnv=0.8)
        pi = tree.pi.copy()
        pi[1::3] = 0.0                                           # zero-mass nodes
        pi /= pi.sum(axis=0, keepdims=True)
        phi = np.zeros_like(pi)
        for k in range(tree.K - 1, -1, -1):                      # parent[k] < k in these generators
            phi[k] += pi[k]
            if tree.parent[k] >= 0:
                phi[tree.parent[k]] += phi[k]
        t2 = O.Tree(tree.parent, tree.node_of_datum, phi, pi)
        rows = np.arange(data.D, dtype=np.int32)
        for sem in (O.SEM_CPP, O.SEM_PY):
            want, got = O.llh_rows(data, t2, rows, sem), host_rows(data, t2, rows, sem)
            assert rel(got, want) < 1e-9, (seed, sem)


@pytest.mark.parametrize("sem", [O.SEM_CPP, O.SEM_PY])
@pytest.mark.parametrize("tiny", [1e-18, 3e-17, 1e-12, 0.0])
def test_states_decided_on_sums_that_cancel(sem, tiny):
    """the decisions "nv == 0" / "nv > 0" / "nr + nv > 0" (mh.hpp:100, params.py:160-166, data.py:52) were taken on the
    sparse signed sums.  SSM 0 sits in node 1 (pi = tiny), its CNV below it in node 2 with copies (2, 0): state 1's variant
    copies come from node 1 alone, nv1 = tiny * 1.  As phisub(1) - phisub(2) that is (0.5 + tiny) - 0.5 = 0 exactly for
    tiny < 1 ulp -- the state would be dropped (Python) or substituted (C++) where the reference keeps it."""
    T = 1
    a = np.array([[30], [25], [900]], dtype=np.int32)
    d = np.array([[60], [50], [1000]], dtype=np.int32)
    mu_r, mu_v = np.array([0.999, 0.999, 0.999]), np.array([0.499, 0.499, 0.5])
    data = O.Data(a, d, mu_r, mu_v, 2, np.array([0, 1, 1], dtype=np.int32), np.array([0], dtype=np.int32),
                  np.array([2], dtype=np.int32), np.array([0], dtype=np.int32))
    parent = np.array([-1, 0, 1], dtype=np.int32)
    pi = np.array([[0.5 - tiny], [tiny], [0.5]])
    phi = np.array([[1.0], [0.5 + tiny], [0.5]])
    nod = np.array([1, 2, 2], dtype=np.int32)
    tree = O.Tree(parent, nod, phi, pi)
    rows = np.arange(3, dtype=np.int32)
    want, got = O.llh_rows(data, tree, rows, sem), host_rows(data, tree, rows, sem)
    assert rel(got, want) < 1e-9, (got, want)
