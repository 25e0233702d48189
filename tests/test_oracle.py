"""-m "not gpu": pins the CPU oracle (oracle/pwgs_oracle.c) against
  (1) tests/golden/ref_vectors.json — outputs of the UNMODIFIED reference mh.cpp/util.cpp (tests/golden/make_golden.py),
  (2) the golden values recorded in SURVEY.md section 4 (same reference build, earlier session),
  (3) the compiled reference itself (oracle/_ref) on fresh random inputs, where it is present,
  (4) published known answers (Philox4x32-10 KAT vectors of Random123)."""
import hashlib
import json
import os
import tempfile

import numpy as np
import pytest

from tests import cases

GOLD = json.load(open(os.path.join(cases.GOLDEN, "ref_vectors.json")))
BY_NAME = {c["name"]: c for c in GOLD["cases"]}


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def load_case(oracle, name):
    if name.startswith("bundled"):
        return cases.bundled(oracle, name.endswith("cnv"))
    return cases.synthetic_case(oracle, **cases.GOLDEN_SYNTH[int(name.split("_")[1])])


def digest(data, tree):
    h = hashlib.sha256()
    for arr in (data.a, data.d, data.mu_r, data.mu_v, data.cnv_ptr, data.cnv_idx, data.cnv_c1, data.cnv_c2,
                tree.parent, tree.node_of_datum, tree.phi, tree.pi):
        h.update(np.ascontiguousarray(arr).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("name", sorted(BY_NAME))
def test_oracle_matches_reference_vectors(oracle, name):
    g = BY_NAME[name]
    data, tree = load_case(oracle, name)
    assert digest(data, tree) == g["sha256"], "seeded generator drifted: regenerate tests/golden/ref_vectors.json"
    assert rel(oracle.multi_param_post(data, tree), g["multi_param_post"]) < 1e-13
    lbc = oracle.lbc_table(data)
    ssm_of_m, coef = oracle.data_states(data, tree)
    m_of = {int(j): m for m, j in enumerate(ssm_of_m)}
    L = oracle.lib()
    import ctypes as C
    L.orc_datum_ll_cpp.restype = C.c_double
    ds = data.cstruct()
    for j, tp, phi, want in g["datum_ll"]:
        cm = coef[m_of[j]].ctypes.data_as(C.POINTER(C.c_int)) if j in m_of else None
        got = L.orc_datum_ll_cpp(C.byref(ds), tree.K, data.T, int(j), int(tp), C.c_double(phi),
                                 tree.pi.ctypes.data_as(C.POINTER(C.c_double)), cm, lbc.ctypes.data_as(C.POINTER(C.c_double)))
        assert rel(got, want) < 1e-13, (j, tp)


@pytest.mark.parametrize("name", [n for n in sorted(BY_NAME) if "mh" in BY_NAME[n]])
def test_oracle_mh_loop_replays_reference_draw_for_draw(oracle, name):
    """Same MT19937 stand-in, same visiting order => the restated mh_loop must land on the reference's final state exactly."""
    g = BY_NAME[name]["mh"]
    data, tree = load_case(oracle, name)
    t2 = cases.children_first(tree)
    got = oracle.mh_loop(data, t2, g["iters"], g["mh_std"], oracle.RNG_MT, recompute_old=True)
    assert got["ratio"] == pytest.approx(g["ratio"], abs=1e-12)
    assert rel(got["pi"], g["pi"]) < 1e-12
    assert rel(got["phi"], g["phi"]) < 1e-12
    # caching the current-state likelihood (what the GPU kernel does) must not change the chain
    got2 = oracle.mh_loop(data, t2, g["iters"], g["mh_std"], oracle.RNG_MT, recompute_old=False)
    assert got2["accepted"] == got["accepted"] and np.array_equal(got2["pi"], got["pi"])


def test_survey_golden_values(oracle):
    # SURVEY.md section 4, Cases A and B
    data, tree = cases.bundled(oracle, False)
    assert rel(oracle.multi_param_post(data, tree), -1872240.7168269958) < 1e-14
    assert rel(oracle.log_bin_coeff(126755, 66023), 87743.31054373167) < 1e-14
    data, tree = cases.bundled(oracle, True)
    assert rel(oracle.multi_param_post(data, tree), -1240573.4871084371) < 1e-14
    ssm, coef = oracle.data_states(data, tree)
    assert list(ssm) == [2, 4]
    # c_data_states lines of Case B: [state][node] = (nr, nv)
    s2 = coef[0]
    assert s2[:, 0].tolist() == [[2, 0]] * 4
    assert s2[:, 1].tolist() == [[1, 2], [2, 1], [2, 1], [2, 1]]
    assert s2[:, 2].tolist() == [[1, 2], [2, 1], [2, 1], [2, 1]]
    s4 = coef[1]
    assert s4[:, 0].tolist() == [[2, 0]] * 4 and s4[:, 1].tolist() == [[1, 0]] * 4 and s4[:, 2].tolist() == [[0, 1]] * 4


def test_scalar_helpers_match_reference_vectors(oracle):
    import ctypes as C
    for n, k, want in GOLD["log_bin_coeff"]:
        assert rel(oracle.log_bin_coeff(n, k), want) < 1e-15 or abs(oracle.log_bin_coeff(n, k) - want) < 1e-9
    for xs, want in GOLD["logsumexp"]:
        arr = (C.c_double * len(xs))(*xs)
        assert oracle.lib().orc_logsumexp(arr, len(xs)) == want


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 with 10 rounds
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = oracle.lib().orc_u01_from_u64(0)
    assert 0 < u < 1e-15 and 0 < oracle.lib().orc_u01_from_u64(2 ** 64 - 1) < 1.0


def test_semantics_agree_where_they_must_and_differ_where_survey_says(oracle):
    """data.py and mh.hpp coincide for plain data and for the bundled CNV example; copy-neutral LOH (2,0) in the same
    node splits them: C++ {st2:1/2, st3:1/2} vs Python {st2:1/3, st3:2/3}."""
    data, tree = cases.bundled(oracle, True)
    assert rel(oracle.total_llh(data, tree, oracle.SEM_PY), oracle.total_llh(data, tree, oracle.SEM_CPP)) < 1e-12
    a = np.array([[30], [70], [500]]); d = np.array([[100], [100], [1000]])
    D = oracle.Data(a, d, [0.999, 0.999, 0.999], [0.5, 0.5, 0.5], 2, [0, 1, 1], [0], [2], [0])
    T = oracle.Tree([-1, 0], [1, 1, 1], [[1.0], [0.6]], [[0.4], [0.6]])
    rows_cpp = oracle.llh_rows(D, T, [0], oracle.SEM_CPP)[0, 1]
    rows_py = oracle.llh_rows(D, T, [0], oracle.SEM_PY)[0, 1]
    lbc = oracle.log_bin_coeff(100, 30)

    def term(nr, nv):
        mu = (nr * 0.999 + nv * 0.001) / (nr + nv)
        return 30 * np.log(mu) + 70 * np.log(1 - mu) + lbc
    st2 = term(0.4 * 2 + 0.6 * 0, 0.6 * 2)      # (cp, cm) = (0, 2): nr = 2*pi_root, nv = 2*pi_1
    st3 = term(0.4 * 2 + 0.6 * 1, 0.6 * 1)      # (t-1, 1)
    assert rows_cpp == pytest.approx(np.logaddexp(st2 + np.log(0.5), st3 + np.log(0.5)), rel=1e-12)
    assert rows_py == pytest.approx(np.logaddexp(st2 + np.log(1 / 3), st3 + np.log(2 / 3)), rel=1e-12)
    assert abs(rows_cpp - rows_py) > 1e-3


def test_all_zero_variant_states(oracle):
    # cm = cp = 0 in the same node: C++ binomial at mu = mu_r (nr+nv>0 via other nodes); Python log(1e-99)  (8a row 11 ii)
    a = np.array([[99], [500]]); d = np.array([[100], [1000]])
    D = oracle.Data(a, d, [0.999, 0.999], [0.5, 0.5], 1, [0, 1], [0], [0], [0])
    T = oracle.Tree([-1], [0, 0], [[1.0]], [[1.0]])
    assert oracle.llh_rows(D, T, [0], oracle.SEM_PY)[0, 0] == pytest.approx(np.log(1e-99))
    # nr+nv == 0 in every state: four log(1e-99) terms (mh.hpp:100) through logsumexp
    assert oracle.llh_rows(D, T, [0], oracle.SEM_CPP)[0, 0] == pytest.approx(np.log(1e-99) + np.log(4.0))


@pytest.mark.parametrize("seed", [101, 102, 103])
def test_oracle_vs_compiled_reference_on_random_inputs(oracle, seed):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference on this box and no prebuilt copy)")
    rng = np.random.default_rng(seed)
    data, tree = cases.synthetic_case(oracle, n_ssm=int(rng.integers(50, 400)), n_cnv=int(rng.integers(2, 20)), T=int(rng.integers(1, 4)),
                                      K=int(rng.integers(2, 15)), seed=seed)
    with tempfile.TemporaryDirectory() as td:
        ref = oracle.RefState(data, tree, td)
        assert rel(oracle.multi_param_post(data, tree), ref.multi_param_post()) < 1e-13
        for j in list(data.cnv_affected[:20]) + [0, data.D - 1]:
            s = tree.node_of_datum[j]
            got = oracle.llh_rows(data, tree, [j], oracle.SEM_CPP)[0, s]
            want = sum(ref.datum_ll(int(j), float(tree.phi[s, tp]), tp) for tp in range(data.T))
            assert rel(got, want) < 1e-12
        ref.close()
