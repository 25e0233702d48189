"""-m gpu: the CUDA library (through the C ABI) against the reference's OWN Python code.

tests/golden/py_vectors.json holds outputs of /root/reference/{data,params,tssb}.py executed under Python 3 (oracle/ref_py3.py,
tests/golden/make_py_golden.py).  Here: K2 (`pwgs_llh_rows`, data.py convention) against `Datum._log_likelihood`, K0b
(`pwgs_get_data_states`) against `params.write_data_state` bit for bit, `pwgs_total_llh` against the data term of
`TSSB.complete_data_log_likelihood`, and the product's assignment sweep driven by the device grid against the reference's
`TSSB.resample_assignments` fed the same uniforms: assignment indices bit-exact.  SURVEY.md 8a rows 10-14."""
import hashlib

import numpy as np
import pytest

from tests import cases, refgold
from tests.test_py_golden import CASES, SWEEPS, LLH_RTOL, case_arrays, check_sweep, digest, oracle_objects, relerr, states_table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=sorted(CASES))
def gcase(request, oracle, gpu_lib):
    g = CASES[request.param]
    data, tree = oracle_objects(oracle, case_arrays(oracle, request.param))
    eng = cases.engine_for(gpu_lib, data, tree)
    yield g, data, tree, eng
    eng.close()


def test_gpu_grid_matches_reference_python_likelihood(gcase, gpu_lib):
    g, data, tree, eng = gcase
    rows = np.array(g["rows"], dtype=np.int32)
    want = np.array(g["grid"])
    got = eng.llh_rows(rows, gpu_lib.SEM_PY)
    assert relerr(got, want) < LLH_RTOL
    sentinel = np.log(1e-99) * data.T
    assert np.array_equal(np.abs(got - sentinel) < 1e-6, np.abs(want - sentinel) < 1e-6)
    got_t = eng.llh_rows(rows, gpu_lib.SEM_PY, col_major=True)
    assert np.array_equal(got_t.T, got)
    cols = np.arange(tree.K - 1, -1, -2, dtype=np.int32)
    assert np.array_equal(eng.llh_block(rows, cols, gpu_lib.SEM_PY), got[:, cols])


def test_gpu_total_llh_matches_reference_data_term(gcase, gpu_lib):
    g, data, tree, eng = gcase
    assert relerr(eng.total_llh(gpu_lib.SEM_PY), g["data_term"]) < LLH_RTOL


def test_gpu_data_states_match_params_write_data_state(gcase):
    g, data, tree, eng = gcase
    ssm_of_m, coef = eng.data_states()
    table = states_table(ssm_of_m, coef)
    assert len(table) == g["data_states_lines"]
    if table:
        assert hashlib.sha256(np.array(sorted(table), dtype=np.int32).tobytes()).hexdigest() == g["data_states_sha256"]
    lookup = {(r[0], r[1]): r[2:] for r in table}
    for line in g["data_states"]:
        assert lookup[(line[0], line[1])] == line[2:], line


@pytest.mark.parametrize("name", sorted(SWEEPS))
@pytest.mark.parametrize("native", [False, True], ids=["python", "native"])
def test_gpu_assignment_sweep_reproduces_the_reference_sweep(oracle, gpu_lib, name, native, capfd):
    from phylowgs_b200.host import model as M
    from phylowgs_b200.host import sampler as S
    from phylowgs_b200.host.backend import GpuBackend
    g = SWEEPS[name]
    arr = case_arrays(oracle, name)
    assert digest(arr) == g["sha256"]
    data = refgold.make_data(M.Datum, arr)
    main, stick = refgold.sticks_for(arr["parent"], g["sticks_seed"])
    backend = GpuBackend(data)
    try:
        tssb, _ = refgold.build_tssb(M.TSSB, M.alleles, data, arr["parent"], arr["node_of_datum"], arr["phi"], arr["pi"], main, stick,
                                     rng=np.random.RandomState(0), backend=backend)
        tssb.rng = refgold.ArrayStream(np.random.RandomState(g["uniform_seed"]).rand(g["n_uniforms"]))
        (S.resample_assignments_native if native else S.resample_assignments_py)(tssb)
        check_sweep(tssb, g)
        assert capfd.readouterr().err.count("Slice sampler shrank") == g["shrank"]
    finally:
        backend.close()
