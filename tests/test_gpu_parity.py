"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Tolerances: integer / index work bit-exact; fp64 log-likelihoods within 1e-9 relative (north_star)."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) if a.size else 0.0


CASES = [
    dict(n_ssm=300, n_cnv=12, T=1, K=6, seed=11),
    dict(n_ssm=500, n_cnv=25, T=3, K=12, seed=12),
    dict(n_ssm=64, n_cnv=6, T=5, K=2, seed=13),
    dict(n_ssm=1000, n_cnv=0, T=2, K=9, seed=14),
    dict(n_ssm=700, n_cnv=40, T=1, K=33, seed=15),
]


@pytest.fixture(scope="module", params=range(len(CASES)))
def case(request, oracle, gpu_lib):
    data, tree = cases.synthetic_case(oracle, **CASES[request.param])
    eng = cases.engine_for(gpu_lib, data, tree)
    yield data, tree, eng
    eng.close()


def test_lbc_matches_log_bin_coeff(case, oracle):
    data, tree, eng = case
    assert rel(eng.lbc(), oracle.lbc_table(data)) < 1e-12


def test_data_states_bit_exact(case, oracle):
    data, tree, eng = case
    ssm_o, coef_o = oracle.data_states(data, tree)
    ssm_g, coef_g = eng.data_states()
    assert np.array_equal(ssm_o, ssm_g)
    assert np.array_equal(coef_o, coef_g)


def test_total_llh_mh_semantics(case, oracle, gpu_lib):
    data, tree, eng = case
    want = oracle.multi_param_post(data, tree)
    assert rel(eng.total_llh(gpu_lib.SEM_MH), want) < RTOL


def test_total_llh_py_semantics(case, oracle, gpu_lib):
    data, tree, eng = case
    want = oracle.total_llh(data, tree, oracle.SEM_PY)
    assert rel(eng.total_llh(gpu_lib.SEM_PY), want) < RTOL


@pytest.mark.parametrize("sem", [0, 1])
def test_llh_rows(case, oracle, sem):
    data, tree, eng = case
    rng = np.random.default_rng(3)
    rows = np.unique(np.concatenate([rng.integers(0, data.D, size=60), data.cnv_affected[:40], np.arange(data.n_ssm, data.D)]))
    want = oracle.llh_rows(data, tree, rows, sem)
    got = eng.llh_rows(rows, sem)
    assert got.shape == want.shape
    assert rel(got, want) < RTOL
    # column-major output (PWGS_LAYOUT_COLS): the same numbers, one contiguous column per node
    got_t = eng.llh_rows(rows, sem, col_major=True)
    assert got_t.shape == (tree.K, len(rows)) and np.array_equal(got_t, got.T)
    cols = np.array([tree.K - 1, 0], dtype=np.int32)
    assert np.array_equal(eng.llh_block(rows[:7], cols, sem, col_major=True), eng.llh_block(rows[:7], cols, sem).T)


def test_bundled_golden_cases(oracle, gpu_lib):
    # SURVEY.md section 4, produced by the unmodified reference param_post
    for with_cnv, golden in ((False, -1872240.7168269958), (True, -1240573.4871084371)):
        data, tree = cases.bundled(oracle, with_cnv)
        with cases.engine_for(gpu_lib, data, tree) as eng:
            assert rel(eng.total_llh(gpu_lib.SEM_MH), golden) < RTOL
            assert rel(eng.total_llh(gpu_lib.SEM_PY), oracle.total_llh(data, tree, oracle.SEM_PY)) < RTOL


@pytest.mark.parametrize("batch", [1, 4, 32, 128, 256])
@pytest.mark.parametrize("ctas", [1, 3, 0])
@pytest.mark.parametrize("stage", [1, 0])
@pytest.mark.parametrize("dist", [-1, 1])
def test_mh_trajectory_matches_oracle(oracle, gpu_lib, batch, ctas, stage, dist):
    """Same Philox stream => the kernel must walk the oracle's chain: equal accept count, equal final state."""
    data, tree = cases.synthetic_case(oracle, n_ssm=400, n_cnv=10, T=2, K=7, seed=21)   # random start: many early accepts
    iters, std, seed, stream = 300, 1000.0, 1234, 7
    want = oracle.mh_loop(data, tree, iters, std, oracle.RNG_PHILOX, seed, stream)
    with cases.engine_for(gpu_lib, data, tree) as eng:
        eng.set_option("mh_batch", batch)
        eng.set_option("mh_ctas", ctas)
        eng.set_option("mh_stage", stage)
        eng.set_option("mh_dist", dist)      # 1: proposals two batches ahead come from the owner CTAs through the global ring
        got = eng.mh_run(iters, std, seed, stream)
    assert got["ratio"] * iters == want["accepted"]
    assert 0 < want["accepted"] < iters
    assert rel(got["pi"], want["pi"]) < 1e-9
    assert rel(got["phi"], want["phi"]) < 1e-9
    assert rel(got["llh"], want["llh"]) < RTOL


def test_mh_zero_iterations_is_identity(oracle, gpu_lib):
    data, tree = cases.bundled(oracle, True)
    with cases.engine_for(gpu_lib, data, tree) as eng:
        got = eng.mh_run(0, 100.0)
    assert np.array_equal(got["pi"], tree.pi) and np.array_equal(got["phi"], tree.phi)
    assert got["ratio"] == 0.0


def test_mh_bundled_example_matches_oracle(oracle, gpu_lib):
    data, tree = cases.bundled(oracle, True)
    want = oracle.mh_loop(data, tree, 500, 100.0, oracle.RNG_PHILOX, 99, 3)
    with cases.engine_for(gpu_lib, data, tree) as eng:
        got = eng.mh_run(500, 100.0, 99, 3)
    assert got["ratio"] * 500 == want["accepted"]
    assert rel(got["pi"], want["pi"]) < 1e-9
    assert rel(got["llh"], want["llh"]) < RTOL


@pytest.mark.parametrize("ctas", [1, 0])
def test_mh_collapsed_plain_data_walks_the_same_chain(oracle, gpu_lib, ctas):
    """Opt-in sufficient statistics (mh_collapse): exact integer totals per (node, class, sample) replace the per-datum terms
    of the plain data; only the order of fp64 additions changes, so the chain must still be the oracle's."""
    data, tree = cases.synthetic_case(oracle, n_ssm=400, n_cnv=10, T=2, K=7, seed=21)
    iters, std, seed, stream = 300, 1000.0, 1234, 7
    want = oracle.mh_loop(data, tree, iters, std, oracle.RNG_PHILOX, seed, stream)
    with cases.engine_for(gpu_lib, data, tree) as eng:
        assert eng.get_option("n_classes") >= 1
        eng.set_option("mh_collapse", 1)
        eng.set_option("mh_ctas", ctas)
        assert rel(eng.total_llh(gpu_lib.SEM_MH), oracle.multi_param_post(data, tree)) < RTOL
        got = eng.mh_run(iters, std, seed, stream)
    assert got["ratio"] * iters == want["accepted"]
    assert rel(got["pi"], want["pi"]) < 1e-9
    assert rel(got["llh"], want["llh"]) < RTOL


@pytest.mark.parametrize("batch", [8, 64, 128, 256])
@pytest.mark.parametrize("ctas", [0, 40])
def test_mh_trajectory_many_rounds_per_cta(oracle, gpu_lib, batch, ctas):
    """Enough data that every CTA holds several warp rounds of both lists, so the cost-levelled partition is exercised with its
    fractional plain rounds (one, two and three proposal-group classes for batch caps 8 / 64 / 128) and ragged tails
    (neither list is a multiple of 32).  The chain must still be the oracle's chain."""
    data, tree = cases.synthetic_case(oracle, n_ssm=20011, n_cnv=60, T=2, K=9, seed=41)
    iters, std, seed, stream = 150, 30000.0, 99, 3
    want = oracle.mh_loop(data, tree, iters, std, oracle.RNG_PHILOX, seed, stream)
    with cases.engine_for(gpu_lib, data, tree) as eng:
        eng.set_option("mh_batch", batch)
        eng.set_option("mh_ctas", ctas)
        got = eng.mh_run(iters, std, seed, stream)
    assert got["ratio"] * iters == pytest.approx(want["accepted"], abs=1e-9)
    assert 0 < want["accepted"] < iters
    assert rel(got["pi"], want["pi"]) < 1e-9
    assert rel(got["phi"], want["phi"]) < 1e-9
    assert rel(got["llh"], want["llh"]) < RTOL


@pytest.mark.parametrize("K,T,iters,std", [(40, 1, 200, 2000.0), (35, 3, 120, 3000.0), (7, 2, 1, 1000.0), (2, 1, 150, 50.0)])
def test_mh_trajectory_wide_trees_and_edge_sizes(oracle, gpu_lib, K, T, iters, std):
    """More nodes than lanes in a warp (the proposal warp strides over nodes), several samples, a single iteration, a
    two-node tree."""
    data, tree = cases.synthetic_case(oracle, n_ssm=350, n_cnv=8, T=T, K=K, seed=30 + K)
    want = oracle.mh_loop(data, tree, iters, std, oracle.RNG_PHILOX, 77, 5)
    with cases.engine_for(gpu_lib, data, tree) as eng:
        got = eng.mh_run(iters, std, 77, 5)
    assert got["ratio"] * iters == pytest.approx(want["accepted"], abs=1e-9)
    assert rel(got["pi"], want["pi"]) < 1e-9
    assert rel(got["phi"], want["phi"]) < 1e-9
    assert rel(got["llh"], want["llh"]) < RTOL


def test_device_log_exp_div_accuracy(gpu_lib):
    """The MH kernel's own branch-free fp64 log / exp / division (csrc/pwgs_device.cuh) against numpy.  The logarithm and the
    exponential are table-driven with short polynomials (measured 2e-14 / 3e-15 relative), the division stops one correction
    short of correct rounding on purpose (2^-40 reciprocal): the bars here are 1e-12 / 1e-13 / 2e-12 relative, three orders of
    magnitude inside the 1e-9 of the likelihood."""
    import ctypes as C
    from phylowgs_b200 import _lib
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(1e-6, 1.0, 20000), rng.uniform(1.0, 700.0, 20000), 10.0 ** rng.uniform(-300, 300, 20000),
                        np.array([1.0, 0.5, 2.0, 0.999, 0.001, 0.499, 0.7071067811865476, 1.4142135623730951, np.nextafter(1.0, 0), np.nextafter(1.0, 2)])])
    out = np.zeros(3 * len(x))
    _lib.check(_lib.load().pwgs_math_selftest(0, len(x), _lib.p_f64(x), _lib.p_f64(out)))
    lg, ex, dv = out[:len(x)], out[len(x):2 * len(x)], out[2 * len(x):]
    want = np.log(x)
    err = np.abs(lg - want) / np.maximum(np.abs(want), 1e-300)
    err[want == 0.0] = np.abs(lg[want == 0.0])
    assert err.max() <= 1e-12, (err.max(), x[err.argmax()])
    small = x <= 700.0
    want_e = np.exp(-x[small])
    err_e = np.abs(ex[small] - want_e) / want_e
    assert err_e.max() <= 1e-13, (err_e.max(), x[small][err_e.argmax()])
    big = x < 1e150
    want_d = x[big] / (1.0 + x[big])
    assert (np.abs(dv[big] - want_d) / want_d).max() <= 2e-12
    print("log max rel err %.2e, exp %.2e, div %.2e" % (err.max(), err_e.max(), (np.abs(dv[big] - want_d) / want_d).max()))
