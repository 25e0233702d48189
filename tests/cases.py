"""Shared seeded inputs for the parity tests (oracle and CUDA path see the same arrays)."""
import os

import numpy as np

from phylowgs_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bundled(oracle, with_cnv):
    """Config 1: the reference's bundled 11-SSM x 5-sample example (+ its single CNV)."""
    data = oracle.read_ssm_cnv(os.path.join(GOLDEN, "ssm_data.txt"), os.path.join(GOLDEN, "cnv_data.txt") if with_cnv else None)
    parent = np.array([-1, 0, 1], dtype=np.int32)
    if with_cnv:   # golden Case B
        nod = np.zeros(12, dtype=np.int32)
        nod[[0, 1, 2, 3, 5, 6, 11]] = 1
        nod[[4, 7, 8, 9, 10]] = 2
    else:          # golden Case A
        nod = np.array([1] * 6 + [2] * 5, dtype=np.int32)
    phi = np.array([[1.0] * 5, [0.8] * 5, [0.3] * 5])
    pi = np.array([[0.2] * 5, [0.5] * 5, [0.3] * 5])
    return data, oracle.Tree(parent, nod, phi, pi)


def random_tree(rng, K, T, D, n_ssm, data=None):
    """Random topology (parent[k] < k), Dirichlet pi, consistent phi, random assignments (root stays empty)."""
    parent = np.array([-1] + [int(rng.integers(0, k)) for k in range(1, K)], dtype=np.int32)
    pi = rng.dirichlet(np.ones(K) * 1.5, size=T).T.copy()
    phi = pi.copy()
    for k in range(K - 1, 0, -1):
        phi[parent[k]] += phi[k]
    nod = rng.integers(1, K, size=D).astype(np.int32) if K > 1 else np.zeros(D, dtype=np.int32)
    return parent, nod, phi, pi


def synthetic_case(oracle, n_ssm, n_cnv, T, K, seed, frac_cnv=0.4, multi_cnv=True, shuffle_nodes=True):
    """Synthetic reads on the truth tree, then a RANDOM tree/assignment so every branch of the state logic
    (unrelated / nested / same-node CNVs, SSMs in several CNVs, zero-copy states) gets exercised."""
    s = synth.make_synthetic(n_ssm, n_cnv, T, 6, seed=seed, frac_cnv=frac_cnv, depth=40)
    rng = np.random.default_rng(seed + 7)
    kw = synth.data_kwargs(s)
    if multi_cnv and n_cnv > 1:
        # add a second CNV link (with awkward copy numbers) to some SSMs
        ptr = kw["cnv_ptr"]
        links = [list(zip(kw["cnv_idx"][ptr[j]:ptr[j + 1]], kw["cnv_c1"][ptr[j]:ptr[j + 1]], kw["cnv_c2"][ptr[j]:ptr[j + 1]]))
                 for j in range(n_ssm)]
        awkward = [(0, 0), (0, 1), (1, 0), (0, 2), (2, 0), (1, 1), (5, 2)]
        for j in range(n_ssm):
            if links[j] and rng.random() < 0.3:
                c = int(rng.integers(0, n_cnv))
                if c != links[j][0][0]:
                    links[j].append((c,) + awkward[int(rng.integers(0, len(awkward)))])
            elif links[j] and rng.random() < 0.3:
                links[j][0] = (links[j][0][0],) + awkward[int(rng.integers(0, len(awkward)))]
        ptr = np.zeros(n_ssm + 1, dtype=np.int32)
        for j in range(n_ssm):
            ptr[j + 1] = ptr[j] + len(links[j])
        flat = [l for ls in links for l in ls]
        kw.update(cnv_ptr=ptr, cnv_idx=np.array([l[0] for l in flat], dtype=np.int32),
                  cnv_c1=np.array([l[1] for l in flat], dtype=np.int32), cnv_c2=np.array([l[2] for l in flat], dtype=np.int32))
    data = oracle.Data(**kw)
    if shuffle_nodes:
        parent, nod, phi, pi = random_tree(rng, K, T, data.D, n_ssm)
    else:
        parent, nod, phi, pi = s["parent"], s["node_of_datum"], s["phi"], s["pi"]
    return data, oracle.Tree(parent, nod, phi, pi)


def engine_for(P, data, tree=None, device=0):
    e = P.Engine(data.a, data.d, data.mu_r, data.mu_v, data.n_ssm, data.cnv_ptr, data.cnv_idx, data.cnv_c1, data.cnv_c2, device=device)
    if tree is not None:
        e.set_tree(tree.parent, tree.node_of_datum, tree.phi, tree.pi)
    return e


# Seeded synthetic cases whose reference outputs are frozen in tests/golden/ref_vectors.json (make_golden.py).
GOLDEN_SYNTH = [
    dict(n_ssm=300, n_cnv=12, T=1, K=6, seed=11),
    dict(n_ssm=500, n_cnv=25, T=3, K=12, seed=12),
    dict(n_ssm=64, n_cnv=6, T=5, K=2, seed=13),
    dict(n_ssm=1000, n_cnv=0, T=2, K=9, seed=14),
    dict(n_ssm=700, n_cnv=40, T=1, K=33, seed=15),
]


def children_first(tree):
    """Relabel nodes k -> K-1-k.  For trees with parent[k] < k (all generators here) index order then lists children before
    parents, which is the file order mh.cpp:134-141 needs; the reference then visits nodes in plain index order and its
    sequential RNG draws line up with the oracle's."""
    from oracle import oracle as O
    K = tree.K
    assert all(tree.parent[k] < k for k in range(K))
    perm = K - 1 - np.arange(K)
    parent = np.full(K, -1, dtype=np.int32)
    for k in range(K):
        parent[perm[k]] = perm[tree.parent[k]] if tree.parent[k] >= 0 else -1
    phi = np.zeros_like(tree.phi)
    pi = np.zeros_like(tree.pi)
    phi[perm] = tree.phi
    pi[perm] = tree.pi
    return O.Tree(parent, perm[tree.node_of_datum].astype(np.int32), phi, pi)
