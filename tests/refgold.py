"""TEST INFRASTRUCTURE shared by tests/golden/make_py_golden.py (which drives the reference's own Python classes, exec'd under
Python 3 by oracle/ref_py3.py) and by the tests that hold the oracle, the host samplers and the CUDA library against the frozen
results.  Everything here is duck-typed on the class names the reference pickles (TSSB / alleles / Datum), so the very same
builder makes a reference tree and a tree of phylowgs_b200.host.model."""
import numpy as np

EPS = np.finfo(np.float64).eps


def edge_case():
    """Hand-built 2-sample data set that walks the branches where data.py and mh.hpp part ways:
    copy-neutral LOH (2,0) / (0,2) in the SSM's own node, (0,0), SSMs above / below / beside their CNV, SSMs in two CNVs,
    a zero-mass node hosting a CNV.  -> dict of arrays in the layout of oracle.Data / oracle.Tree."""
    rng = np.random.default_rng(4242)
    n_ssm, n_cnv, T = 16, 4, 2
    links = {1: [(0, 2, 0)], 2: [(0, 0, 2)], 3: [(0, 0, 0)], 4: [(1, 2, 1)], 5: [(1, 1, 2)], 6: [(0, 2, 1), (1, 0, 2)],
             7: [(2, 1, 1), (1, 3, 0)], 8: [(3, 2, 0)], 9: [(3, 1, 0)], 10: [(2, 0, 1)], 11: [(0, 1, 0)], 13: [(0, 2, 2), (2, 0, 2)],
             14: [(1, 0, 0), (3, 1, 1)], 15: [(2, 2, 0)]}
    parent = np.array([-1, 0, 1, 1, 3, 1], dtype=np.int32)
    node_of_datum = np.array([1, 1, 1, 1, 1, 4, 3, 1, 4, 3, 5, 5, 2, 1, 3, 2] + [1, 3, 2, 4], dtype=np.int32)
    pi = np.array([[0.1, 0.2], [0.3, 0.1], [0.2, 0.25], [0.15, 0.3], [0.0, 0.0], [0.25, 0.15]])
    phi = pi.copy()
    for k in range(len(parent) - 1, 0, -1):
        phi[parent[k]] += phi[k]
    d = rng.integers(40, 160, size=(n_ssm + n_cnv, T)).astype(np.int32)
    a = (d * rng.uniform(0.45, 0.98, size=d.shape)).astype(np.int32)
    d[n_ssm:] *= 40
    a[n_ssm:] = (d[n_ssm:] * 0.7).astype(np.int32)
    ptr = np.zeros(n_ssm + 1, dtype=np.int32)
    idx, c1, c2 = [], [], []
    for j in range(n_ssm):
        for (c, x, y) in links.get(j, []):
            idx.append(c), c1.append(x), c2.append(y)
        ptr[j + 1] = len(idx)
    return dict(a=a, d=d, mu_r=np.full(n_ssm + n_cnv, 0.999), mu_v=np.array([0.499] * n_ssm + [0.5] * n_cnv), n_ssm=n_ssm,
                cnv_ptr=ptr, cnv_idx=np.array(idx, dtype=np.int32), cnv_c1=np.array(c1, dtype=np.int32),
                cnv_c2=np.array(c2, dtype=np.int32), parent=parent, node_of_datum=node_of_datum, phi=phi, pi=pi)


def sticks_for(parent, seed):
    """Seeded `main` per node and the stick of every node inside its parent's stick array (root: main 1e-30 and, for its first
    child, .999999 — what evolve.py:84-98 / tssb.py:179,187 keep them at)."""
    rng = np.random.RandomState(seed)
    K = len(parent)
    main = rng.uniform(0.05, 0.6, size=K)
    stick = rng.uniform(0.2, 0.8, size=K)
    main[0] = 1e-30
    for k in range(1, K):
        if parent[k] == 0:
            stick[k] = 0.999999
    return main, stick


def make_data(Datum, arr):
    """Datum objects (SSMs s0.. then CNVs c0..) with the (cnv, c1, c2) links of util2.py:86-90."""
    n_ssm, D = int(arr["n_ssm"]), len(arr["a"])
    out = []
    for j in range(D):
        ident = ("s%d" % j) if j < n_ssm else ("c%d" % (j - n_ssm))
        out.append(Datum(ident if j < n_ssm else ident, ident, [int(x) for x in arr["a"][j]], [int(x) for x in arr["d"][j]],
                         float(arr["mu_r"][j]), float(arr["mu_v"][j])))
    for j in range(n_ssm):
        for l in range(int(arr["cnv_ptr"][j]), int(arr["cnv_ptr"][j + 1])):
            out[j].cnv.append((out[n_ssm + int(arr["cnv_idx"][l])], int(arr["cnv_c1"][l]), int(arr["cnv_c2"][l])))
    return out


def build_tssb(TSSB, alleles, data, parent, node_of_datum, phi, pi, main, stick, hypers=(25.0, 1.0, 0.25), **tssb_kw):
    """A TSSB holding the given tree: node k's children in increasing index order.  -> (tssb, [node object of index k]).
    Works for the reference's classes and for phylowgs_b200.host.model's (same constructor arguments, same tree dicts)."""
    K, T = len(parent), len(data[0].a)
    root = alleles(conc=0.1, ntps=T)
    tssb = TSSB(dp_alpha=hypers[0], dp_gamma=hypers[1], alpha_decay=hypers[2], root_node=root, data=data, **tssb_kw)
    tssb.root["main"] = float(main[0])
    entries, nodes = {0: tssb.root}, {0: root}
    kids = [[c for c in range(K) if parent[c] == k] for k in range(K)]
    order = [0]
    for k in order:
        for c in kids[k]:
            node = nodes[k].spawn()
            entry = {"node": node, "main": float(main[c]), "sticks": np.empty((0, 1)), "children": []}
            entries[k]["sticks"] = np.vstack([entries[k]["sticks"], float(stick[c])])
            entries[k]["children"].append(entry)
            entries[c], nodes[c] = entry, node
            order.append(c)
    for k in range(K):
        nodes[k].pi = np.array(pi[k], dtype=float)
        nodes[k].params = np.array(phi[k], dtype=float)
        nodes[k].id = k
    for n in range(len(data)):
        tssb.assignments[n].remove_datum(n)
        nodes[int(node_of_datum[n])].add_datum(n)
        tssb.assignments[n] = nodes[int(node_of_datum[n])]
        data[n].node = nodes[int(node_of_datum[n])]
    for dat in data:
        dat.tssb = tssb
    return tssb, [nodes[k] for k in range(K)]


def node_paths(tssb):
    """{id(node): tuple of child indices from the root} for every node of the tree."""
    out, stack = {}, [(tssb.root, ())]
    while stack:
        entry, path = stack.pop()
        out[id(entry["node"])] = path
        for i, c in enumerate(entry["children"]):
            stack.append((c, path + (i,)))
    return out


def tree_summary(tssb):
    """What a sweep leaves behind, in a form that does not depend on object identity: per datum the child-index path of its
    node; per node that holds data (by path) main, the sticks of its children, pi and params; totals over all nodes."""
    paths = node_paths(tssb)
    assign = ["/".join(map(str, paths[id(nd)])) for nd in tssb.assignments]
    nodes, stack = {}, [(tssb.root, ())]
    n_nodes, sum_main, sum_stick, sum_pi = 0, 0.0, 0.0, 0.0
    while stack:
        entry, path = stack.pop()
        n_nodes += 1
        sum_main += float(entry["main"])
        sum_stick += float(np.sum(entry["sticks"]))
        sum_pi += float(np.sum(entry["node"].pi))
        if len(entry["node"].data) or len(path) <= 1:       # full detail for the nodes that hold data (and the top two levels)
            nodes["/".join(map(str, path))] = dict(main=float(entry["main"]), sticks=[float(x) for x in np.ravel(entry["sticks"])],
                                                   pi=[float(x) for x in np.ravel(entry["node"].pi)],
                                                   params=[float(x) for x in np.ravel(entry["node"].params)],
                                                   n_data=len(entry["node"].data))
        for i, c in enumerate(entry["children"]):
            stack.append((c, path + (i,)))
    return dict(assign=assign, nodes=nodes, n_nodes=n_nodes, sum_main=sum_main, sum_stick=sum_stick, sum_pi=sum_pi)


class ArrayStream(object):
    """The draw source both sides of the sweep comparison consume: a fixed array of uniforms handed out in order."""

    def __init__(self, uniforms):
        self.u = np.asarray(uniforms, dtype=float)
        self.pos = 0

    def next(self):
        v = float(self.u[self.pos])
        self.pos += 1
        return v

    # numpy.random.RandomState face used by phylowgs_b200.host.sampler.UniformStream (rng.rand(block))
    def rand(self, n=None):
        if n is None:
            return self.next()
        n = int(n)
        out = np.zeros(n)
        m = min(n, len(self.u) - self.pos)
        out[:m] = self.u[self.pos:self.pos + m]
        self.pos += m                                          # a block may run past the end: the rest is never consumed
        return out


def boundbeta1(u, b):
    """boundbeta(1, b) by inversion from one uniform — the form the product's sweep uses (host/model.py), substituted for
    util.boundbeta on the reference side too so that both sides consume the same uniforms."""
    return (1.0 - EPS) * ((1.0 - u ** (1.0 / b)) - 0.5) + 0.5
