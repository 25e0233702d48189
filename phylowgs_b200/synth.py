"""Synthetic tumours of the shapes BASELINE.json names.  Pure numpy, deterministic
in `seed`; used by bench.py and the parity tests.  Values follow the conventions of the reference's
input files (README.md:21-55, parser/create_phylowgs_inputs.py:311-345,461-475): a = reference reads,
d = total reads, mu_r = 0.999, mu_v = 0.499, CNV rows "sK,minor,major"."""
import os

import numpy as np

CN_STATES = [(2, 1), (2, 0), (1, 0), (2, 2), (3, 1)]   # (major, minor)


def truth_tree(k_true, rng, T, n_empty=2):
    """Empty normal root (node 0) + k_true cancerous nodes, mixed linear/branching, + n_empty empty leaves.
    Node indices are a preorder numbering only by construction of `parent` (parent[k] < k)."""
    base = [-1, 0, 1, 1, 2, 2, 3, 4, 6]
    parent = [base[k] if k < len(base) else int(rng.integers(1, k)) for k in range(k_true + 1)]
    K0 = len(parent)
    for e in range(n_empty):
        parent.append(int(rng.integers(1, K0)))
    parent = np.array(parent, dtype=np.int32)
    K = len(parent)
    kids = [[c for c in range(K) if parent[c] == k] for k in range(K)]
    phi = np.zeros((K, T))
    pi = np.zeros((K, T))
    phi[0] = 1.0
    for tp in range(T):
        def split(n):
            ch = kids[n]
            if not ch:
                pi[n, tp] = phi[n, tp]
                return
            real = [c for c in ch if c < K0]
            w = rng.dirichlet(np.ones(len(real) + 1) * 2.0) if real else np.array([1.0])
            if n == 0:
                w = np.array([0.15, 0.85])          # clonal population at phi = 0.85
            own = phi[n, tp] * w[0]
            for c, f in zip(real, w[1:]):
                phi[c, tp] = phi[n, tp] * f
            for c in ch:
                if c >= K0:                          # empty extra node: a sliver of its parent's own mass
                    phi[c, tp] = own * 1e-3
                    own -= phi[c, tp]
            pi[n, tp] = own
            for c in ch:
                split(c)
        split(0)
    return parent, phi, pi, K0


def make_synthetic(n_ssm=50000, n_cnv=300, T=1, k_true=8, seed=1003, frac_cnv=0.3, depth=60, n_empty=2):
    """-> dict(a, d [D,T], mu_r, mu_v [D], n_ssm, cnv_ptr, cnv_idx, cnv_c1, cnv_c2, parent, node_of_datum, phi, pi)."""
    rng = np.random.default_rng(seed)
    parent, phi, pi, K0 = truth_tree(k_true, rng, T, n_empty)
    K = len(parent)
    D = n_ssm + n_cnv
    canc = np.arange(1, K0)
    ssm_node = rng.choice(canc, size=n_ssm, p=rng.dirichlet(np.ones(len(canc)) * 2.0)).astype(np.int32)
    cnv_node = rng.choice(canc, size=n_cnv).astype(np.int32) if n_cnv else np.zeros(0, dtype=np.int32)
    node_of_datum = np.concatenate([ssm_node, cnv_node]).astype(np.int32)
    mu_r = np.full(D, 0.999)
    mu_v = np.concatenate([np.full(n_ssm, 0.499), np.full(n_cnv, 0.5)])
    a = np.zeros((D, T), dtype=np.int32)
    d = np.zeros((D, T), dtype=np.int32)

    # ---- SSM -> CNV links: a fraction of SSMs falls in exactly one CNV
    in_cnv = (rng.random(n_ssm) < frac_cnv) if n_cnv else np.zeros(n_ssm, dtype=bool)
    link_cnv = np.where(in_cnv, rng.integers(0, max(n_cnv, 1), size=n_ssm), -1)
    state = np.array(CN_STATES)[rng.integers(0, len(CN_STATES), size=max(n_cnv, 1))]   # per CNV (major, minor)
    cnv_ptr = np.concatenate([[0], np.cumsum(in_cnv)]).astype(np.int32)
    cnv_idx = link_cnv[in_cnv].astype(np.int32)
    cnv_c1 = state[cnv_idx, 1].astype(np.int32) if in_cnv.any() else np.zeros(0, dtype=np.int32)   # "sK,minor,major"
    cnv_c2 = state[cnv_idx, 0].astype(np.int32) if in_cnv.any() else np.zeros(0, dtype=np.int32)

    # ---- ancestor matrix anc[u, v] = u is ancestor-or-self of v
    anc = np.zeros((K, K), dtype=bool)
    for v in range(K):
        u = v
        while u >= 0:
            anc[u, v] = True
            u = parent[u]

    # ---- plain SSM reads
    plain = ~in_cnv
    for tp in range(T):
        dd = np.maximum(rng.poisson(depth, size=n_ssm), 5)
        mu = (1 - phi[ssm_node, tp]) * 0.999 + phi[ssm_node, tp] * 0.499
        aa = rng.binomial(dd, mu)
        d[:n_ssm, tp] = dd
        a[:n_ssm, tp] = aa
    # ---- CNV-affected SSM reads from the true (nr, nv) genome mixture (the four cases of data.py:71-109)
    idx = np.nonzero(in_cnv)[0]
    if len(idx):
        s = ssm_node[idx][:, None]                       # [M,1]
        cn = cnv_node[link_cnv[idx]][:, None]
        c1 = cnv_c1[:, None].astype(float)
        c2 = cnv_c2[:, None].astype(float)
        t = c1 + c2
        ks = np.arange(K)[None, :]
        in_sub = anc[s, ks]
        under = anc[cn, ks]
        s_above_cnv = anc[s[:, 0], cn[:, 0]][:, None]
        early = rng.random((len(idx), 1)) < 0.5          # variant allele on the amplified (c2) copy or not
        nr_d = np.where(s_above_cnv, np.where(early, c1, c2), np.maximum(0, t - 1))
        nv_d = np.where(s_above_cnv, np.where(early, c2, c1), np.minimum(1, t))
        cr = np.where(~in_sub & ~under, 2.0, np.where(in_sub & ~under, 1.0, np.where(~in_sub & under, t, nr_d)))
        cv = np.where(~in_sub & ~under, 0.0, np.where(in_sub & ~under, 1.0, np.where(~in_sub & under, 0.0, nv_d)))
        for tp in range(T):
            nr = cr @ pi[:, tp]
            nv = cv @ pi[:, tp]
            tot = np.maximum(nr + nv, 1e-9)
            mu = (nr * 0.999 + nv * 0.001) / tot
            dd = np.maximum(rng.poisson(depth * tot / 2.0), 5)
            d[idx, tp] = dd
            a[idx, tp] = rng.binomial(dd, np.clip(mu, 0, 1))
    # ---- CNV data rows
    for tp in range(T):
        if n_cnv:
            dd = rng.integers(1000, 90001, size=n_cnv)
            d[n_ssm:, tp] = dd
            a[n_ssm:, tp] = ((1 - phi[cnv_node, tp] / 2.0) * dd).astype(np.int32)
    return dict(a=a, d=d, mu_r=mu_r, mu_v=mu_v, n_ssm=n_ssm, cnv_ptr=cnv_ptr, cnv_idx=cnv_idx, cnv_c1=cnv_c1, cnv_c2=cnv_c2,
                parent=parent, node_of_datum=node_of_datum, phi=phi, pi=pi)


def config(n):
    """The BASELINE.json configs that are synthetic (3: WGS tumour, 4: multi-region); seed = 1000 + n."""
    if n == 3:
        return make_synthetic(50000, 300, 1, 8, seed=1003, frac_cnv=0.3, depth=60)
    if n == 4:
        return make_synthetic(20000, 0, 5, 8, seed=1004, frac_cnv=0.0, depth=100)
    raise ValueError("config %r is not synthetic" % (n,))


def from_files(ssm_fn, cnv_fn, k_pops):
    """Bench state for a data set in the reference's file formats (README.md:21-55): a chain tree root -> 1 -> ... -> k_pops, the
    SSMs dealt to the populations by their variant-allele fraction in the first sample (equal-size bins), every CNV in
    population 1, phi of a population = twice the mean VAF of its SSMs (made non-increasing down the chain, at most 1)."""
    from .host.backend import arrays_from_data
    from .host.evolve import load_data
    codes, n_ssm, n_cnv, _ = load_data(ssm_fn, cnv_fn if cnv_fn else os.devnull)
    arr = arrays_from_data(codes)
    a, d = arr["a"], arr["d"]
    D, T = a.shape
    K = k_pops + 1
    vaf = 1.0 - a[:n_ssm] / np.maximum(d[:n_ssm], 1)
    order = np.argsort(-vaf[:, 0], kind="stable")
    node = np.ones(D, dtype=np.int32)
    for r, j in enumerate(order):
        node[j] = 1 + min(k_pops - 1, r * k_pops // max(n_ssm, 1))
    phi = np.ones((K, T))
    for k in range(1, K):
        m = vaf[node[:n_ssm] == k].mean(axis=0) if np.any(node[:n_ssm] == k) else phi[k - 1] / 2
        phi[k] = np.minimum(np.clip(2.0 * m, 0.02, 1.0), phi[k - 1] * 0.98)
    pi = phi.copy()
    pi[:-1] -= phi[1:]
    parent = np.arange(-1, K - 1, dtype=np.int32)
    arr.update(parent=parent, node_of_datum=node, phi=phi, pi=pi)
    return arr


def data_kwargs(s):
    return dict(a=s["a"], d=s["d"], mu_r=s["mu_r"], mu_v=s["mu_v"], n_ssm=s["n_ssm"], cnv_ptr=s["cnv_ptr"],
                cnv_idx=s["cnv_idx"], cnv_c1=s["cnv_c1"], cnv_c2=s["cnv_c2"])
