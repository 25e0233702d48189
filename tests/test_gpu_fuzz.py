"""-m gpu: randomised shapes and kernel options against the oracle's chain (same Philox stream => same accepts, same final state).
PWGS_FUZZ_CASES raises the number of cases for soak runs (scripts/gpu_soak.sh)."""
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
N_CASES = int(os.environ.get("PWGS_FUZZ_CASES", "40"))


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) if a.size else 0.0


def test_random_shapes_and_options_walk_the_oracle_chain(oracle, gpu_lib):
    rng = np.random.default_rng(int(os.environ.get("PWGS_FUZZ_SEED", "2026")))
    seen_accepts = 0
    for case in range(N_CASES):
        n_ssm = int(rng.choice([40, 200, 700, 3000, 9000]))
        n_cnv = int(rng.choice([0, 0, 3, 12, 40])) if n_ssm >= 200 else int(rng.choice([0, 2]))
        T = int(rng.integers(1, 6))
        K = int(rng.choice([2, 3, 5, 9, 17, 33, 40, 64, 100]))      # 64, 100: node tables in global memory, Euler-tour scan
        data, tree = cases.synthetic_case(oracle, n_ssm=n_ssm, n_cnv=n_cnv, T=T, K=K, seed=1000 + case)
        iters = int(rng.choice([1, 7, 60, 250]))
        std = float(rng.choice([30.0, 300.0, 3000.0, 30000.0]))
        seed, stream = int(rng.integers(1, 2 ** 40)), int(rng.integers(0, 1000))
        opts = dict(mh_batch=int(rng.choice([1, 2, 8, 16, 64, 128, 256])), mh_ctas=int(rng.choice([0, 0, 1, 2, 7, 32, 60, 148])),
                    mh_stage=int(rng.integers(0, 2)), mh_dist=int(rng.choice([-1, -1, 0, 1])), mh_collapse=int(rng.integers(0, 2)))
        want = oracle.mh_loop(data, tree, iters, std, oracle.RNG_PHILOX, seed, stream)
        with cases.engine_for(gpu_lib, data, tree) as eng:
            if eng.get_option("n_classes") == 0:
                opts["mh_collapse"] = 0
            for k, v in opts.items():
                eng.set_option(k, v)
            got = eng.mh_run(iters, std, seed, stream)
        what = "case %d: n_ssm=%d n_cnv=%d T=%d K=%d iters=%d std=%g %r" % (case, n_ssm, n_cnv, T, K, iters, std, opts)
        assert round(got["ratio"] * iters) == want["accepted"], what
        assert rel(got["pi"], want["pi"]) < 1e-9, what
        assert rel(got["phi"], want["phi"]) < 1e-9, what
        assert rel(got["llh"], want["llh"]) < 1e-9, what
        seen_accepts += want["accepted"]
    assert seen_accepts > 0
