"""Assign further SSMs to the trees of a finished run (misc/post_assign_ssm.py:11-216), the second consumer of the datum x node
likelihood grid.

    python -m phylowgs_b200.host.post_assign_ssm [--cnvs cnv_data.txt] [--seed N] ssm_data.txt trees.zip s17 s42 ...

For every tree of trees.zip (burn-in first, then the samples, util2.py:325-331, empty vertices removed, util2.py:128-155) and every
requested SSM: the SSM is added as one more datum, starts in a node drawn from where its former node-mates sit now
(post_assign_ssm.py:84-91), and one slice-sampling step over the nodes -- weighted by their share of the data
(post_assign_ssm.py:106-111) -- picks its node.  Prints {"assignments": {ssm id: [node index per tree]}, "trees": path} with node
indices in the descending-phi order of post_assign_ssm.py:113-124.

The reference evaluates `node.logprob([datum])` per candidate (O(N) Python work each, data.py:33-38); here the whole row -- the
new datum against all nodes of the tree -- is one device call per tree (`Backend.llh_rows`, K2), the slice sampler only indexes it.
Randomness: a numpy RandomState (seedable) drawing in the reference's order; its global generator cannot be replayed (Python 2)."""
import argparse
import collections
import csv
import json
import os
import pickle
import sys
import zipfile

import numpy as np

from .evolve import install_pickle_aliases
from .model import Datum


# ------------------------------------------------------------------ trees.zip reader (util2.py:264-343)
def remove_empty_nodes(entry, parent=None):
    """Drop nodes without data; an empty inner node hands its children (and their sticks) to its parent (util2.py:128-155).
    The root is kept."""
    for child in list(entry["children"]):
        remove_empty_nodes(child, entry)
    if entry["node"].data or parent is None:
        return
    at = parent["children"].index(entry)
    if entry["children"]:
        node_parent = entry["node"].parent()
        for i, child in enumerate(list(entry["children"])):
            parent["children"].append(child)
            parent["sticks"] = np.append(parent["sticks"], np.reshape(entry["sticks"][i], (1, 1)), 0)
            entry["children"].remove(child)
        for child in list(entry["node"].children()):
            child._parent = node_parent
            node_parent.add_child(child)
            entry["node"].remove_child(child)
    parent["children"].remove(entry)
    parent["sticks"] = np.delete(parent["sticks"], at, 0)
    entry["node"].kill()


def read_trees(trees_fn, remove_empty_vertices=True):
    """-> (index, tree) for the burn-in trees (negative indices, ascending) and then the samples (util2.py:325-331)."""
    install_pickle_aliases()
    with zipfile.ZipFile(trees_fn) as z:
        names = z.namelist()
        order = sorted((int(n.split("_")[1]), n) for n in names if n.startswith("burnin_")) + \
            sorted((int(n.split("_")[1]), n) for n in names if n.startswith("tree_"))
        for idx, name in order:
            tree = pickle.loads(z.read(name))
            if remove_empty_vertices:
                remove_empty_nodes(tree.root)
            yield idx, tree


# ------------------------------------------------------------------ inputs (post_assign_ssm.py:133-166)
def read_requested_ssms(ssm_fn, cnv_fn, ssm_ids):
    """-> [(name, id, a, d, mu_r, mu_v, [(cnv id, c1, c2), ...])] for the requested ids, in file order."""
    links = collections.defaultdict(list)
    if cnv_fn:
        with open(cnv_fn, newline="") as f:
            for row in csv.DictReader(f, delimiter="\t"):
                for tok in filter(None, (row.get("ssms") or "").split(";")):
                    sid, c1, c2 = tok.split(",")
                    links[sid].append((row["cnv"], int(c1), int(c2)))
    want, out = set(ssm_ids), []
    with open(ssm_fn, newline="") as f:
        for row in csv.DictReader(f, delimiter="\t"):
            if row["id"] in want:
                out.append((row["gene"], row["id"], [int(x) for x in row["a"].split(",")], [int(x) for x in row["d"].split(",")],
                            float(row["mu_r"]), float(row["mu_v"]), links.get(row["id"], [])))
    return out


# ------------------------------------------------------------------ the sampler
def index_map(tree, nodes):
    """nodes.index -> position in a depth-first walk that visits children by descending mean phi (post_assign_ssm.py:113-124)."""
    pos = {id(nd): i for i, nd in enumerate(nodes)}
    mapping, stack = {}, [tree.root["node"]]
    while stack:
        nd = stack.pop()
        mapping[pos[id(nd)]] = len(mapping)
        stack.extend(sorted(nd.children(), key=lambda c: np.mean(c.params)))          # popped last-first: descending
    return mapping


def starting_node(tree, nodes, mates, rng):
    """A node drawn by where the SSM's node-mates of the previous tree are assigned now (post_assign_ssm.py:84-91); the first
    non-root node when there are none."""
    if not mates:
        return nodes[1]
    counts = collections.OrderedDict()
    for i in mates:
        nd = tree.assignments[i]
        counts[id(nd)] = (nd, counts.get(id(nd), (nd, 0))[1] + 1)
    shares = np.cumsum([c / float(len(mates)) for _, c in counts.values()])
    rng.rand()                                            # the reference draws one uniform it does not use (line 89)
    return list(counts.values())[min(int(np.sum(rng.rand() > shares)), len(counts) - 1)][0]


def post_assignments(ssms, trees_fn, make_backend, rng=None):
    """make_backend(data) -> Backend over `data` = the trees' data with the new SSMs appended to the SSM block (GpuBackend in the
    product).  One device context for the whole job: an SSM's likelihood does not look at other SSMs, so all requested SSMs ride
    along in the same data set and a tree costs one tree upload plus one grid call of len(ssms) rows."""
    rng = rng or np.random.RandomState()
    eps = np.finfo(np.float64).eps
    assignments, mates_of = collections.defaultdict(list), collections.defaultdict(list)
    backend = perm = new_rows = None
    for _, tree in read_trees(trees_fn):
        n = len(tree.data)
        nodes = tree.get_nodes()
        pos = {id(nd): k for k, nd in enumerate(nodes)}
        if backend is None:
            data, perm, new_rows = device_data(tree.data, ssms)
            backend = make_backend(data)
        # the tree as arrays; the new SSMs sit in node 1 for the upload (a row is "if it sat in node k", whatever its own node)
        parent = np.array([-1 if nd.parent() is None else pos[id(nd.parent())] for nd in nodes], dtype=np.int32)
        nod = np.ones(n + len(ssms), dtype=np.int32)
        nod[perm] = [pos[id(nd)] for nd in tree.assignments[:n]]
        phi = np.array([np.asarray(nd.params, dtype=np.float64).reshape(-1) for nd in nodes])
        pi = np.array([np.asarray(nd.pi, dtype=np.float64).reshape(-1) for nd in nodes])
        backend.set_tree_arrays(parent, nod, phi, pi)
        rows = backend.llh_rows(new_rows)                 # [len(ssms), K]
        sizes = np.array([nd.num_local_data() for nd in nodes], dtype=float)
        mapping = index_map(tree, nodes)
        for row, (name, sid, _a, _d, _mr, _mv, _links) in zip(rows, ssms):
            cur = pos[id(starting_node(tree, nodes, mates_of[name], rng))]
            with_new = sizes.copy()
            with_new[cur] += 1.0                          # the datum sits in its current node while candidates are drawn
            shares = np.cumsum(with_new / float(n))
            llh_s = np.log(rng.rand()) + row[cur]
            lo, hi = 0.0, 1.0
            while True:
                u = (hi - lo) * rng.rand() + lo
                cand = min(int(np.sum(u > shares)), len(nodes) - 1)
                if row[cand] > llh_s:
                    cur = cand
                    break
                if abs(hi - lo) < eps:
                    break                                 # slice sampler shrank down: keep the current node
                if cur > cand:
                    lo = u
                else:
                    hi = u
            mates_of[name] = list(nodes[cur].data)
            assignments[sid].append(mapping[cur])
    if backend is not None and hasattr(backend, "close"):
        backend.close()
    return dict(assignments)


def device_data(base, ssms):
    """-> (data for the device: the trees' SSMs, then the new SSMs, then the CNVs, re-labelled s0.. / c0.. as pwgs_create wants;
    perm[i] = device row of tree.data[i]; rows of the new SSMs).  tree.data has the same order in every tree of a run."""
    ssm_pos = [i for i, x in enumerate(base) if x.id[0] == "s"]
    cnv_pos = [i for i, x in enumerate(base) if x.id[0] != "s"]
    n_old, n_new = len(ssm_pos), len(ssms)
    perm = np.zeros(len(base), dtype=np.int64)
    perm[ssm_pos] = np.arange(n_old)
    perm[cnv_pos] = n_old + n_new + np.arange(len(cnv_pos))
    out = [None] * (len(base) + n_new)
    for i, x in enumerate(base):
        row = int(perm[i])
        out[row] = Datum(x.name, ("s%d" % row) if x.id[0] == "s" else ("c%d" % (row - n_old - n_new)), x.a, x.d, x.mu_r, x.mu_v)
    by_name = {x.name: out[int(perm[i])] for i, x in enumerate(base)}
    for i, x in enumerate(base):
        out[int(perm[i])].cnv = [(by_name[c.name], c1, c2) for c, c1, c2 in x.cnv]
    for k, (name, _sid, a, d, mu_r, mu_v, links) in enumerate(ssms):
        y = Datum(name, "s%d" % (n_old + k), a, d, mu_r, mu_v)
        y.cnv = [(by_name[c], c1, c2) for c, c1, c2 in links if c in by_name]       # post_assign_ssm.py:26-29
        out[n_old + k] = y
    return out, perm, np.arange(n_old, n_old + n_new, dtype=np.int32)


def gpu_backend_factory(device=0):
    from .backend import GpuBackend
    return lambda data: GpuBackend(data, device=device)


def main(argv=None):
    p = argparse.ArgumentParser(description="Assign mutations to existing trees", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--cnvs", dest="cnv_file", help="Path to CNV list")
    p.add_argument("--seed", type=int, default=None, help="Seed of the sampler's random generator")
    p.add_argument("ssm_file", help="SSM file whose mutations you wish to reassign")
    p.add_argument("trees_file", help="Path to sampled trees")
    p.add_argument("ssm_ids", nargs="+", help="ID of SSM in ssm_file to reassign")
    args = p.parse_args(argv)
    ssms = read_requested_ssms(args.ssm_file, args.cnv_file, args.ssm_ids)
    result = post_assignments(ssms, args.trees_file, gpu_backend_factory(int(os.environ.get("PWGS_DEVICE", "0"))),
                              np.random.RandomState(args.seed))
    sys.stdout.write(json.dumps({"assignments": result, "trees": os.path.realpath(args.trees_file)}) + "\n")


if __name__ == "__main__":
    main()
