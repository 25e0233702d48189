"""-m gpu: the MH kernel at the shapes BASELINE.json names (configs 3 and 4), trees whose node tables do not fit in shared memory
(mh_kernel<false, true>: tables in global memory) and several chains in one cooperative launch (pwgs_mh_launch_multi).
Tolerances as in test_gpu_parity.py: 1e-9 relative on fp64 log-likelihoods / states, accept counts exact."""
import numpy as np
import pytest

from phylowgs_b200 import synth
from tests import cases
from tests.test_gpu_parity import RTOL, rel

pytestmark = pytest.mark.gpu


def _oracle_objects(oracle, s):
    data = oracle.Data(**synth.data_kwargs(s))
    tree = oracle.Tree(s["parent"], s["node_of_datum"], s["phi"], s["pi"])
    return data, tree


@pytest.mark.parametrize("config,iters,std", [(3, 150, 1.0e6), (3, 120, 1.0e3), (4, 120, 1.0e3), (4, 120, 1.0e5)])
def test_named_config_shapes(oracle, gpu_lib, config, iters, std):
    """BASELINE.json configs[2] (50k SSMs, 300 CNVs, T=1) and configs[3] (20k SSMs x 5 samples): both likelihood conventions at
    the bench state (truth tree), then a trajectory with the default launch (batch 256, all SMs) against the oracle's chain.
    At the truth the chain of config 4 accepts nothing at any proposal width (the reference's +1e-4 perturbation, util.cpp:26-27,
    costs more log-likelihood over 100 000 terms than any move gains), so the trajectories there, and one of config 3's, start
    from a random tree and assignment over the same data: dozens of accepts, every batch size of the ramp."""
    s = synth.config(config)
    data, tree = _oracle_objects(oracle, s)
    with cases.engine_for(gpu_lib, data, tree) as eng:
        assert rel(eng.total_llh(gpu_lib.SEM_MH), oracle.multi_param_post(data, tree)) < RTOL
        assert rel(eng.total_llh(gpu_lib.SEM_PY), oracle.total_llh(data, tree, oracle.SEM_PY)) < RTOL
        if std < 1.0e6:
            parent, nod, phi, pi = cases.random_tree(np.random.default_rng(5), tree.K, data.T, data.D, data.n_ssm)
            tree = oracle.Tree(parent, nod, phi, pi)
            eng.set_tree(tree.parent, tree.node_of_datum, tree.phi, tree.pi)
            assert rel(eng.total_llh(gpu_lib.SEM_MH), oracle.multi_param_post(data, tree)) < RTOL
        want = oracle.mh_loop(data, tree, iters, std, oracle.RNG_PHILOX, 2026, 11)
        got = eng.mh_run(iters, std, 2026, 11)
    assert got["ratio"] * iters == pytest.approx(want["accepted"], abs=1e-9)
    assert 0 < want["accepted"] < iters
    assert rel(got["pi"], want["pi"]) < 1e-9
    assert rel(got["phi"], want["phi"]) < 1e-9
    assert rel(got["llh"], want["llh"]) < RTOL


@pytest.mark.parametrize("batch", [4, 32, 256])
@pytest.mark.parametrize("ctas", [1, 3, 0])
def test_node_tables_in_global_memory_same_chain(oracle, gpu_lib, batch, ctas):
    """mh_big = 1 forces the variant that large trees need onto a small case: it must walk the oracle's chain too."""
    data, tree = cases.synthetic_case(oracle, n_ssm=400, n_cnv=10, T=2, K=7, seed=21)
    iters, std, seed, stream = 300, 1000.0, 1234, 7
    want = oracle.mh_loop(data, tree, iters, std, oracle.RNG_PHILOX, seed, stream)
    with cases.engine_for(gpu_lib, data, tree) as eng:
        eng.set_option("mh_big", 1)
        eng.set_option("mh_batch", batch)
        eng.set_option("mh_ctas", ctas)
        got = eng.mh_run(iters, std, seed, stream)
    assert got["ratio"] * iters == want["accepted"]
    assert 0 < want["accepted"] < iters
    assert rel(got["pi"], want["pi"]) < 1e-9
    assert rel(got["phi"], want["phi"]) < 1e-9
    assert rel(got["llh"], want["llh"]) < RTOL


@pytest.mark.parametrize("K,T", [(160, 4), (400, 5), (1500, 2)])
def test_trees_beyond_the_shared_memory_budget(oracle, gpu_lib, K, T):
    """K*T = 640 .. 3000 node/sample entries (round 1 failed with PWGS_ELIMIT from ~1270): early sweeps with dp_alpha up to 50 or
    many samples reach such trees, and the reference has no limit."""
    data, tree = cases.synthetic_case(oracle, n_ssm=600, n_cnv=12, T=T, K=K, seed=50 + T)
    iters, std = 40, 4000.0 * K
    want = oracle.mh_loop(data, tree, iters, std, oracle.RNG_PHILOX, 5, 9)
    with cases.engine_for(gpu_lib, data, tree) as eng:
        assert rel(eng.total_llh(gpu_lib.SEM_MH), oracle.multi_param_post(data, tree)) < RTOL
        got = eng.mh_run(iters, std, 5, 9)
    assert got["ratio"] * iters == pytest.approx(want["accepted"], abs=1e-9)
    assert rel(got["pi"], want["pi"]) < 1e-9
    assert rel(got["phi"], want["phi"]) < 1e-9
    assert rel(got["llh"], want["llh"]) < RTOL


@pytest.mark.parametrize("n_chains", [2, 4, 8])
def test_chains_in_one_launch_equal_their_own_launches(oracle, gpu_lib, n_chains):
    """pwgs_mh_launch_multi: chain i on n_sms / n_chains CTAs of ONE cooperative grid gives the bits of its own launch with
    mh_ctas = n_sms / n_chains (different data sets, trees, proposal widths and streams per chain)."""
    specs = [dict(n_ssm=300 + 170 * i, n_cnv=6 + i, T=1 + i % 3, K=5 + 2 * i, seed=60 + i) for i in range(n_chains)]
    built = [cases.synthetic_case(oracle, **sp) for sp in specs]
    engs = [cases.engine_for(gpu_lib, d, t) for d, t in built]
    try:
        cpc = engs[0].get_option("n_sms") // n_chains
        iters = 200
        stds = [500.0 * (i + 1) for i in range(n_chains)]
        seeds = [100 + i for i in range(n_chains)]
        streams = [7 * i + 1 for i in range(n_chains)]
        alone = []
        for i, e in enumerate(engs):
            e.set_option("mh_ctas", cpc)
            alone.append(e.mh_run(iters, stds[i], seeds[i], streams[i]))
            e.set_option("mh_ctas", 0)
        gpu_lib.Engine.mh_launch_multi(engs, iters, stds, seeds, streams)
        for i, e in enumerate(engs):
            got = e.mh_wait()
            assert got["ratio"] == alone[i]["ratio"]
            assert np.array_equal(got["pi"], alone[i]["pi"]) and np.array_equal(got["phi"], alone[i]["phi"])
            assert got["llh"] == alone[i]["llh"]
        # and chain 0 against the oracle
        want = oracle.mh_loop(built[0][0], built[0][1], iters, stds[0], oracle.RNG_PHILOX, seeds[0], streams[0])
        assert alone[0]["ratio"] * iters == want["accepted"]
        assert rel(alone[0]["llh"], want["llh"]) < RTOL
    finally:
        for e in engs:
            e.close()


def test_pitched_range_lands_in_the_columns_of_a_wider_grid(oracle, gpu_lib):
    """pwgs_llh_range_pitched: the column-major block of pwgs_llh_range written in place at out[c * pitch + r] (what the native
    sweep uses for the columns of spawned nodes); bytes around the block stay as they were."""
    import ctypes as C
    from phylowgs_b200 import _lib
    data, tree = cases.synthetic_case(oracle, n_ssm=700, n_cnv=12, T=2, K=9, seed=33)
    with cases.engine_for(gpu_lib, data, tree) as eng:
        D, row0, cols = data.D, 123, np.array([7, 2, 5], dtype=np.int32)
        want = eng.llh_range(row0, D - row0, cols, gpu_lib.SEM_PY, col_major=True)
        grid = np.full((5, D), -7.0)
        _lib.check(_lib.load().pwgs_llh_range_pitched(eng._h, row0, D - row0, len(cols), _lib.p_i32(cols), gpu_lib.SEM_PY | _lib.LAYOUT_COLS,
                                                     grid[1:, row0:].ctypes.data_as(C.POINTER(C.c_double)), D))
        assert np.array_equal(grid[1:4, row0:], want)
        assert (grid[0] == -7.0).all() and (grid[4] == -7.0).all() and (grid[1:4, :row0] == -7.0).all()
        rc = _lib.load().pwgs_llh_range_pitched(eng._h, row0, D - row0, len(cols), _lib.p_i32(cols), gpu_lib.SEM_PY,
                                                grid[1:, row0:].ctypes.data_as(C.POINTER(C.c_double)), D)
        assert rc == -1                                                     # row-major output cannot be pitched


@pytest.mark.parametrize("sem", [0, 1])
def test_grid_kernel_with_sorted_rows_gives_the_file_order_bits(oracle, gpu_lib, sem):
    """pair_llh_sorted_kernel (plain data and CNV-affected SSMs of a range in separate warps) against pair_llh_kernel walking the
    same range in file order (option grid_plain): the same function per pair, so the same bits, for whole grids, ranges that
    start inside either kind, node subsets and ranges without any datum of one kind."""
    data, tree = cases.synthetic_case(oracle, n_ssm=900, n_cnv=25, T=2, K=10, seed=44)
    with cases.engine_for(gpu_lib, data, tree) as eng:
        D = data.D
        for row0, n_rows, cols in [(0, D, None), (1, D - 1, None), (457, D - 457, np.array([3, 9, 0], dtype=np.int32)),
                                   (900, 25, None), (880, 40, np.array([5], dtype=np.int32)), (13, 1, None), (D - 1, 1, None)]:
            eng.set_option("grid_plain", 1)
            want = eng.llh_range(row0, n_rows, cols, sem, col_major=True)
            eng.set_option("grid_plain", 0)
            got = eng.llh_range(row0, n_rows, cols, sem, col_major=True)
            assert np.array_equal(want, got), (row0, n_rows)
        rows = np.arange(D, dtype=np.int32)
        assert rel(eng.llh_rows(rows, sem), oracle.llh_rows(data, tree, rows, sem)) < RTOL
        assert np.array_equal(eng.llh_range(0, D, None, sem, col_major=True).T, eng.llh_rows(rows, sem))
