"""The two likelihood-bound TSSB call sites, restructured around batched device calls.

* `resample_assignments` (tssb.py:82-150): the slice sampler over the unit interval is inherently sequential (each datum's
  move changes node occupancy, and `find_node` may spawn nodes), but the only expensive thing it asks for is
  "log-likelihood of datum n if it sat in node k".  The reference answers that with one Python `logprob` call that re-derives
  tree metadata every time (O(N + K) each, O(N^2) per sweep, data.py:33-38).  Here the whole datum x node grid comes from ONE
  device call per tree version (pwgs_llh_rows, Python convention of data.py); it is refreshed only when the tree changes in a
  way that can change unread rows: a node was spawned (new column, and the parent's pi moved), or a CNV datum moved (its SSMs'
  rows).  Decisions compare fp64 numbers produced by the device against host-drawn uniforms, so given the same uniforms the
  assignment indices are those of the CPU oracle.
* `complete_data_log_likelihood` (tssb.py:404-410): sum_k n_k log w_k on the host + one device reduction for the data term.
"""
import sys

import numpy as np

from .model import EPS


def _path_cmp(cur, new):
    """Ordering of two child-index paths in the stick-breaking layout of [0,1) (tssb.py:84-94: the reference compares the
    zero-padded "%03d" strings with py2 cmp(s2, s1); for fewer than 1000 children per node that is the lexicographic order of
    the index tuples).  < 0: the candidate lies left of the current node, so the lower bound moves up; else the upper bound."""
    if not cur and not new:
        return 0
    if not cur:
        return 1
    if not new:
        return -1
    a, b = tuple(cur), tuple(new)
    return (b > a) - (b < a)


class _Grid(object):
    """llh[n, k]: log-likelihood of datum n if it sat in node k, filled by device calls and kept valid through the sweep.

    Nodes get stable indices for the sweep: preorder position at the start, then spawn order (pwgs_set_tree takes any node
    order), so columns never move.  What invalidates entries that are still to be read:
      * `find_node` spawned nodes: only their columns are missing.  A spawn splits its parent's pi and leaves every existing
        node's phi unchanged, and every likelihood depends on the tree only through phi (plain data directly; CNV-affected SSMs
        through the subtree sums their piecewise-constant coefficients reduce to) -> one block call, remaining rows x new nodes;
      * a CNV datum moved: the rows of the SSMs linked to it that are still to come -> one block call, those rows x all nodes;
      * an SSM moved: nothing (no other datum's likelihood looks at it).
    """

    def __init__(self, tssb):
        self.tssb = t = tssb
        self.backend = b = t.backend
        self.nodes = list(b.sync_tree(t))                   # preorder; the device tree now equals the host tree
        self.uid = {id(nd): k for k, nd in enumerate(self.nodes)}
        self.nod = np.zeros(t.num_data, dtype=np.int32)
        for k, nd in enumerate(self.nodes):
            if nd.data:
                self.nod[list(nd.data)] = k
        self.spawned = getattr(t, "n_spawned", 0)
        self.calls = 1
        full = b.llh_rows(np.arange(t.num_data, dtype=np.int32))
        self.llh = np.empty((t.num_data, max(2 * full.shape[1], 32)))
        self.llh[:, :full.shape[1]] = full
        if not hasattr(t, "_ssms_of_cnv"):                  # CNV datum index -> SSM datum indices linked to it
            m = {}
            for d in t.data:
                for cnv, _, _ in d.cnv:
                    m.setdefault(cnv.index, []).append(d.index)
            t._ssms_of_cnv = {k: np.array(sorted(v), dtype=np.int32) for k, v in m.items()}

    def _push_tree(self):
        nodes = self.nodes
        parent = np.array([-1 if nd.parent() is None else self.uid[id(nd.parent())] for nd in nodes], dtype=np.int32)
        phi = np.array([np.asarray(nd.params, dtype=np.float64).reshape(-1) for nd in nodes])
        pi = np.array([np.asarray(nd.pi, dtype=np.float64).reshape(-1) for nd in nodes])
        self.backend.set_tree_arrays(parent, self.nod, phi, pi)

    def after_spawn(self, n):
        """Columns of the nodes spawned since the last look, for the data still to be visited."""
        new = [nd for nd in self.tssb.get_nodes() if id(nd) not in self.uid]
        first = len(self.nodes)
        for nd in new:
            self.uid[id(nd)] = len(self.nodes)
            self.nodes.append(nd)
        if len(self.nodes) > self.llh.shape[1]:
            grown = np.empty((self.llh.shape[0], 2 * len(self.nodes)))
            grown[:, :first] = self.llh[:, :first]
            self.llh = grown
        self._push_tree()
        cols = np.arange(first, len(self.nodes), dtype=np.int32)
        self.llh[n:, first:len(self.nodes)] = self.backend.llh_block(np.arange(n, self.tssb.num_data, dtype=np.int32), cols)
        self.spawned = getattr(self.tssb, "n_spawned", 0)
        self.calls += 1

    def after_cnv_move(self, n, cnv_index):
        rows = self.tssb._ssms_of_cnv.get(cnv_index)
        if rows is None:
            return
        rows = rows[rows > n]
        if len(rows) == 0:
            return
        self._push_tree()
        K = len(self.nodes)
        self.llh[rows, :K] = self.backend.llh_block(rows, np.arange(K, dtype=np.int32))
        self.calls += 1

    def get(self, n, node):
        return self.llh[n, self.uid[id(node)]]


def resample_assignments(tssb):
    rng, root_entry = tssb.rng, tssb.root
    grid = _Grid(tssb)
    n_ssm = tssb.backend.n_ssm
    paths = {}
    for n in range(tssb.num_data):
        datum = tssb.data[n]
        cur = tssb.assignments[n]
        indices = paths.get(id(cur))
        if indices is None:                                 # child-index path of the node (tssb.py:101-107); spawns append
            indices, entry = [], root_entry                 # children at the end, so paths stay valid through the sweep
            for anc in cur.get_ancestors()[1:]:
                k = next(i for i, c in enumerate(entry["children"]) if c["node"] is anc)
                indices.append(k)
                entry = entry["children"][k]
            paths[id(cur)] = indices
        lo, hi = 0.0, 1.0
        llh_s = np.log(rng.rand()) + grid.get(n, cur)
        while True:
            u = (hi - lo) * rng.rand() + lo
            cand, path = tssb.find_node(u)
            if cand.parent() is None:                       # keep the root empty (tssb.py:118-120)
                cand = cand.children()[0]
                path = [0]
            if getattr(tssb, "n_spawned", 0) != grid.spawned:
                grid.after_spawn(n)
            if grid.get(n, cand) > llh_s:
                if cand is not cur:
                    cur.remove_datum(n)
                    cand.add_datum(n)
                    tssb.assignments[n] = cand
                    datum.node = cand
                    grid.nod[n] = grid.uid[id(cand)]
                    if datum.index >= n_ssm:
                        grid.after_cnv_move(n, datum.index)
                break
            if abs(hi - lo) < EPS:
                sys.stderr.write("Slice sampler shrank down.  Keep current state.\n")
                break
            if _path_cmp(indices, path) < 0:
                lo = u
            else:
                hi = u
    tssb.last_grid_refreshes = grid.calls


def complete_data_log_likelihood(tssb):
    weights, nodes = tssb.get_mixture()
    tssb.backend.sync_tree(tssb)
    total = 0.0
    for w, nd in zip(weights, nodes):
        if nd.num_local_data():
            total += nd.num_local_data() * np.log(w)
    return float(total + tssb.backend.total_llh())
