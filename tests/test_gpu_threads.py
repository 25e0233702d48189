"""-m gpu: one context per chain, many chains as threads of one process (multievolve --in-process).  Each thread's results
must be bit-identical to the same calls made alone: contexts share the device, the kernels and their function attributes,
nothing else."""
import threading

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

# different K*T and data sizes -> different shared-memory footprints and batch caps of the MH kernel
SHAPES = [
    dict(n_ssm=400, n_cnv=10, T=1, K=4, seed=31),
    dict(n_ssm=900, n_cnv=30, T=4, K=40, seed=32),
    dict(n_ssm=2000, n_cnv=0, T=2, K=9, seed=33),
    dict(n_ssm=300, n_cnv=20, T=6, K=70, seed=34),
    dict(n_ssm=64, n_cnv=6, T=5, K=2, seed=35),
    dict(n_ssm=1500, n_cnv=60, T=3, K=25, seed=36),
]
ROUNDS = 12


def chain_calls(P, data, tree, rounds):
    """A chain's call pattern: tree upload, MH move, likelihood grid, total likelihood -- with the MH result fed back."""
    eng = cases.engine_for(P, data)
    out = []
    phi, pi = tree.phi.copy(), tree.pi.copy()
    rows = np.arange(0, data.D, max(data.D // 97, 1), dtype=np.int32)
    for r in range(rounds):
        eng.set_tree(tree.parent, tree.node_of_datum, phi, pi)
        res = eng.mh_run(150, 80.0, seed=data.D, stream=r)
        phi, pi = res["phi"], res["pi"]
        eng.set_tree(tree.parent, tree.node_of_datum, phi, pi)
        out.append((res["phi"].copy(), res["pi"].copy(), res["ratio"], res["llh"], eng.llh_rows(rows, P.SEM_PY).copy(),
                    eng.total_llh(P.SEM_PY)))
    eng.close()
    return out


def same(a, b):
    return all(np.array_equal(np.asarray(x), np.asarray(y)) for ra, rb in zip(a, b) for x, y in zip(ra, rb)) and len(a) == len(b)


def test_concurrent_chains_equal_the_same_chains_run_alone(gpu_lib, oracle):
    inputs = [cases.synthetic_case(oracle, **s) for s in SHAPES]
    alone = [chain_calls(gpu_lib, d, t, ROUNDS) for d, t in inputs]
    got, errors = [None] * len(inputs), []

    def body(i):
        try:
            got[i] = chain_calls(gpu_lib, inputs[i][0], inputs[i][1], ROUNDS)
        except Exception as e:                      # noqa: BLE001 - reported below with the chain's index
            errors.append((i, repr(e)))
    for attempt in range(3):
        threads = [threading.Thread(target=body, args=(i,)) for i in range(len(inputs))]
        for t in threads:
            t.start()
        for t in threads:
            t.join(300)
        assert not errors, errors
        for i in range(len(inputs)):
            assert same(got[i], alone[i]), "chain %d differs when run beside the others (attempt %d)" % (i, attempt)
