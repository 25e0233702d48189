"""Hand-made trees.zip archives for the write_results tests: trees of chosen shape over a real data file, pickled by the product's
own writer (host/fastpickle.py) under the member names evolve.py gives them (util2.py:250-262)."""
import json
import zipfile

import numpy as np

from phylowgs_b200.host import evolve as E
from phylowgs_b200.host import fastpickle
from phylowgs_b200.host.model import TSSB, alleles


def build_tree(ssm_fn, cnv_fn, parent, phi, assign, seed=0):
    """parent[k] < k (node 0 = root, holds no data), phi [K][T] consistent (children's sum <= parent's), assign [n_data] in 1..K-1."""
    E.install_pickle_aliases()
    codes, n_ssms, n_cnvs, mapping = E.load_data(ssm_fn, cnv_fn)
    rng = np.random.RandomState(seed)
    T = len(codes[0].a)
    phi = np.asarray(phi, dtype=np.float64).reshape(len(parent), -1)
    if phi.shape[1] == 1 and T > 1:
        phi = phi * np.linspace(1.0, 0.8, T)[None, :]                      # a different prevalence in every sample, still consistent
    root = alleles(conc=0.1, ntps=T)
    t = TSSB(dp_alpha=25.0, dp_gamma=1.0, alpha_decay=0.25, root_node=root, data=codes, rng=rng, backend=None)
    entries = [t.root]
    for k in range(1, len(parent)):
        pe = entries[parent[k]]
        e = {"node": pe["node"].spawn(), "main": float(rng.uniform(0.1, 0.9)), "sticks": np.empty((0, 1)), "children": []}
        pe["children"].append(e)
        pe["sticks"] = np.vstack([pe["sticks"], rng.uniform(0.1, 0.9)])
        entries.append(e)
    for k, e in enumerate(entries):
        kids = [c for c in range(len(parent)) if parent[c] == k]
        e["node"].params = phi[k].copy()
        e["node"].pi = phi[k] - sum((phi[c] for c in kids), np.zeros(T))
        assert (e["node"].pi > -1e-12).all()
    for n in range(t.num_data):
        root.remove_datum(n)
        nd = entries[int(assign[n])]["node"]
        nd.add_datum(n)
        t.assignments[n] = nd
    for d in codes:
        d.tssb = t
    E.map_datum_to_node(t)
    return t, mapping


def write_archive(path, trees, mapping, params=None, burnin=()):
    """trees: [(tssb, llh)] -> members tree_<i>_<llh>; burnin: trees written as burnin_<-n..-1>_<llh>."""
    with zipfile.ZipFile(path, "w", compression=zipfile.ZIP_DEFLATED, allowZip64=True) as z:
        z.writestr("cnv_logical_physical_mapping.json", json.dumps(mapping))
        if params is not None:
            z.writestr("params.json", json.dumps(params))
        for i, (t, llh) in enumerate(burnin):
            z.writestr("burnin_%d_%s" % (i - len(burnin), llh), bytes(fastpickle.dumps(t)))
        for i, (t, llh) in enumerate(trees):
            z.writestr("tree_%d_%s" % (i, llh), bytes(fastpickle.dumps(t)))


def shaped_forest(ssm_fn, cnv_fn, n_data, seed):
    """A few trees per shape the result filters care about: superclones, multiprimary roots, tiny populations, deep chains,
    random branchings.  -> ([(tssb, llh)], mapping)"""
    rng = np.random.RandomState(seed)
    shapes = []
    # clonal population with few SSMs over one child with most of them at nearly the same prevalence: a superclone
    a = np.full(n_data, 2); a[:max(1, n_data // 8)] = 1
    shapes.append(([-1, 0, 1], [1.0, 0.82, 0.79], a))
    # the same, but the prevalences are far apart: kept
    shapes.append(([-1, 0, 1], [1.0, 0.9, 0.4], a))
    # superclone with a grandchild (the clonal node still has exactly one child)
    b = np.where(np.arange(n_data) % 5 == 0, 3, 2); b[:max(1, n_data // 10)] = 1
    shapes.append(([-1, 0, 1, 2], [1.0, 0.8, 0.75, 0.3], b))
    # multiprimary: two children of the root
    shapes.append(([-1, 0, 0, 1], [1.0, 0.55, 0.35, 0.2], 1 + np.arange(n_data) % 3))
    # tiny populations in the middle and at a leaf: their mutations move to the population with the closest implied phi
    c = 1 + np.arange(n_data) % 2; c[0] = 3; c[1] = 4; c[-1] = 5
    shapes.append(([-1, 0, 1, 1, 3, 2], [1.0, 0.9, 0.5, 0.3, 0.1, 0.25], c))
    # a deep chain
    shapes.append(([-1, 0, 1, 2, 3, 4], [1.0, 0.95, 0.7, 0.5, 0.3, 0.1], 1 + np.arange(n_data) % 5))
    # children listed in increasing phi: the numbering must follow decreasing phi
    shapes.append(([-1, 0, 1, 1, 1], [1.0, 0.9, 0.1, 0.2, 0.4], 1 + (np.arange(n_data) * 7) % 4))
    for _ in range(5):                                                     # random branchings
        K = int(rng.randint(3, 9))
        parent = [-1, 0] + [int(rng.randint(1, k)) for k in range(2, K)]
        phi = np.zeros(K); phi[0] = 1.0
        for k in range(1, K):
            sibs = sum(phi[c] for c in range(1, k) if parent[c] == parent[k])
            phi[k] = (phi[parent[k]] - sibs) * rng.uniform(0.3, 0.9)
        w = rng.dirichlet(np.ones(K - 1) * 0.6)
        shapes.append((parent, phi, 1 + rng.choice(K - 1, size=n_data, p=w)))
    trees, mapping = [], None
    for i, (parent, phi, assign) in enumerate(shapes):
        t, mapping = build_tree(ssm_fn, cnv_fn, parent, phi, assign, seed=seed + i)
        trees.append((t, -1000.0 - float(rng.uniform(0, 500))))
    return trees, mapping
