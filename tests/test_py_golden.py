"""-m "not gpu": the Python half of the path held against the reference's OWN Python code.

tests/golden/py_vectors.json was produced by executing /root/reference/{data,params,tssb,alleles,node,util2}.py under Python 3
(oracle/ref_py3.py: declared mechanical patches, nothing restated; generator tests/golden/make_py_golden.py).  Here the CPU oracle
(SEM_PY likelihood, compute_n_genomes, write_data_state, write_tree) and the host samplers of the product (the assignment sweep in
both its forms, complete_data_log_likelihood) are compared with those vectors.  tests/test_gpu_py_golden.py does the same for
the CUDA library.  SURVEY.md 8a rows 9-14."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests import cases, refgold

GOLD = json.load(open(os.path.join(cases.GOLDEN, "py_vectors.json")))
CASES = {c["name"]: c for c in GOLD["cases"]}
SWEEPS = {s["name"]: s for s in GOLD["sweeps"]}

# north_star's tolerance.  The reference adds x*log(mu) + (n-x)*log(1-mu) + log_bin_coeff in Python / numpy order, the oracle in
# C: terms of 1e2..1e5 cancel to values that can be < 1, so the observed agreement is 1e-14 of the terms, ~2e-11 of the result.
LLH_RTOL = 1e-9


def case_arrays(oracle, name):
    if name == "edge":
        return refgold.edge_case()
    if name == "sim_4_100_50":
        data = oracle.read_ssm_cnv(os.path.join(cases.GOLDEN, "emptysim.4.100.50.ssm.txt"))
        parent, nod, phi, pi = cases.random_tree(np.random.default_rng(77), 7, data.T, data.D, data.n_ssm)
        tree = oracle.Tree(parent, nod, phi, pi)
    elif name.startswith("bundled"):
        data, tree = cases.bundled(oracle, name.endswith("cnv"))
    else:
        data, tree = cases.synthetic_case(oracle, **cases.GOLDEN_SYNTH[int(name.split("_")[1])])
    return dict(a=data.a, d=data.d, mu_r=data.mu_r, mu_v=data.mu_v, n_ssm=data.n_ssm, cnv_ptr=data.cnv_ptr, cnv_idx=data.cnv_idx,
                cnv_c1=data.cnv_c1, cnv_c2=data.cnv_c2, parent=tree.parent, node_of_datum=tree.node_of_datum, phi=tree.phi, pi=tree.pi)


def digest(arr):
    h = hashlib.sha256()
    for key in ("a", "d", "mu_r", "mu_v", "cnv_ptr", "cnv_idx", "cnv_c1", "cnv_c2", "parent", "node_of_datum", "phi", "pi"):
        h.update(np.ascontiguousarray(arr[key]).tobytes())
    return h.hexdigest()


def oracle_objects(oracle, arr):
    data = oracle.Data(arr["a"], arr["d"], arr["mu_r"], arr["mu_v"], arr["n_ssm"], arr["cnv_ptr"], arr["cnv_idx"], arr["cnv_c1"], arr["cnv_c2"])
    tree = oracle.Tree(arr["parent"], arr["node_of_datum"], arr["phi"], arr["pi"])
    return data, tree


def relerr(got, want):
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))) if got.size else 0.0


def states_table(ssm_of_m, coef):
    """oracle.data_states output -> sorted rows [ssm, node, nr1, nv1, .., nr4, nv4] (the layout of the fixture)."""
    M, _, K, _ = coef.shape if len(coef) else (0, 4, 0, 2)
    rows = []
    for m in range(M):
        for k in range(K):
            rows.append([int(ssm_of_m[m]), k] + [int(x) for q in range(4) for x in coef[m, q, k]])
    return rows


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_python_likelihood_matches_the_reference_python(oracle, name):
    g = CASES[name]
    arr = case_arrays(oracle, name)
    assert digest(arr) == g["sha256"], "seeded generator drifted: regenerate tests/golden/py_vectors.json"
    data, tree = oracle_objects(oracle, arr)
    rows = np.array(g["rows"], dtype=np.int32)
    got = oracle.llh_rows(data, tree, rows, oracle.SEM_PY)
    want = np.array(g["grid"])
    assert got.shape == want.shape
    assert relerr(got, want) < LLH_RTOL
    # the sentinel log(1e-99) (data.py:56-57) must be hit in exactly the same cells
    sentinel = np.log(1e-99)
    assert np.array_equal(np.abs(got - sentinel * data.T) < 1e-6, np.abs(want - sentinel * data.T) < 1e-6)
    assert relerr(oracle.total_llh(data, tree, oracle.SEM_PY), g["data_term"]) < LLH_RTOL


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_compute_n_genomes_matches_the_reference_python(oracle, name):
    g = CASES[name]
    data, tree = oracle_objects(oracle, case_arrays(oracle, name))
    for j_str, per_node in g["n_genomes"].items():
        j = int(j_str)
        for k, per_tp in enumerate(per_node):
            for tp, states in enumerate(per_tp):
                ns, nr, nv = oracle.compute_n_genomes(data, tree, j, k, tp)
                assert ns == len(states), (j, k, tp)                      # 4 states iff one CNV in the SSM's own node (data.py:122-126)
                want = np.array(states)
                assert np.allclose(nr[:ns], want[:, 0], rtol=1e-13, atol=1e-15), (j, k, tp)
                assert np.allclose(nv[:ns], want[:, 1], rtol=1e-13, atol=1e-15), (j, k, tp)
                # exact zeros decide which states data.py:52 drops
                assert np.array_equal(nv[:ns] == 0, want[:, 1] == 0), (j, k, tp)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_data_states_match_params_write_data_state(oracle, name):
    g = CASES[name]
    data, tree = oracle_objects(oracle, case_arrays(oracle, name))
    ssm_of_m, coef = oracle.data_states(data, tree)
    table = states_table(ssm_of_m, coef)
    assert len(table) == g["data_states_lines"]
    if table:
        assert hashlib.sha256(np.array(sorted(table), dtype=np.int32).tobytes()).hexdigest() == g["data_states_sha256"]
    lookup = {(r[0], r[1]): r[2:] for r in table}
    for line in g["data_states"]:
        assert lookup[(line[0], line[1])] == line[2:], line


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_c_tree_matches_params_write_tree(oracle, name, tmp_path):
    """Same nodes in the same (post-)order, same children, same data ids, same heights, phi / pi to the last bit."""
    g = CASES[name]
    data, tree = oracle_objects(oracle, case_arrays(oracle, name))
    fn = str(tmp_path / "c_tree.txt")
    oracle.write_c_tree(tree, fn)
    got = [ln.rstrip("\n").split("\t") for ln in open(fn)]
    want = [ln.split("\t") for ln in g["c_tree"]]
    assert len(got) == len(want) == tree.K
    for a, b in zip(got, want):
        assert a[0] == b[0] and a[3] == b[3] and a[5] == b[5] and a[7] == b[7], (a, b)          # id, #children, #data, height
        assert a[4].split(",") == b[4].split(",")                                                # children in the same order
        assert sorted(a[6].split(",")) == sorted(b[6].split(","))                                # data ids (a set in the reference)
        assert [float(x) for x in a[1].split(",")] == [float(x) for x in b[1].split(",")]
        assert [float(x) for x in a[2].split(",")] == [float(x) for x in b[2].split(",")]


def host_tree(oracle, name, seed, backend_cls=None):
    from phylowgs_b200.host import model as M
    from tests.oracle_backend import OracleBackend
    arr = case_arrays(oracle, name)
    data = refgold.make_data(M.Datum, arr)
    for i, d in enumerate(data):
        d.index = i
    main, stick = refgold.sticks_for(arr["parent"], seed)
    backend = (backend_cls or OracleBackend)(data)
    tssb, nodes = refgold.build_tssb(M.TSSB, M.alleles, data, arr["parent"], arr["node_of_datum"], arr["phi"], arr["pi"], main, stick,
                                     rng=np.random.RandomState(0), backend=backend)
    return tssb, nodes, arr


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_complete_data_log_likelihood_matches_the_reference(oracle, name):
    g = CASES[name]
    tssb, _, _ = host_tree(oracle, name, g["sticks_seed"])
    assert relerr(tssb.complete_data_log_likelihood(), g["cdll"]) < LLH_RTOL


def check_sweep(tssb, g):
    got = refgold.tree_summary(tssb)
    assert got["n_nodes"] == g["n_nodes"]
    assert got["assign"] == g["assign"]                                   # assignment indices: bit-exact (north_star)
    assert sorted(got["nodes"]) == sorted(g["nodes"])
    for path, w in g["nodes"].items():
        h = got["nodes"][path]
        assert h["n_data"] == w["n_data"], path
        assert h["main"] == w["main"] and h["sticks"] == w["sticks"], path
        assert np.allclose(h["pi"], w["pi"], rtol=1e-15, atol=0) and np.allclose(h["params"], w["params"], rtol=1e-15, atol=0), path
    for key in ("sum_main", "sum_stick", "sum_pi"):
        assert got[key] == pytest.approx(g[key], rel=1e-12)


@pytest.mark.parametrize("name", sorted(SWEEPS))
@pytest.mark.parametrize("native", [False, True], ids=["python", "native"])
def test_assignment_sweep_reproduces_the_reference_sweep(oracle, name, native, capfd):
    """TSSB.resample_assignments of the reference (tssb.py:82-150, run by make_py_golden.py with its draws taken from a fixed
    array of uniforms) against the product's sweep fed the same array: same node for every datum, same spawned nodes with the
    same sticks and the same pi, same number of "shrank down" events."""
    from phylowgs_b200.host import sampler as S
    g = SWEEPS[name]
    tssb, _, arr = host_tree(oracle, name, g["sticks_seed"])
    assert digest(arr) == g["sha256"]
    uniforms = np.random.RandomState(g["uniform_seed"]).rand(g["n_uniforms"])
    tssb.rng = refgold.ArrayStream(uniforms)
    (S.resample_assignments_native if native else S.resample_assignments_py)(tssb)
    check_sweep(tssb, g)
    assert capfd.readouterr().err.count("Slice sampler shrank") == g["shrank"]
