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
    """Ordering of two child-index paths in the stick-breaking layout of [0,1) (tssb.py:84-94: zero-padded strings, py2 cmp).
    < 0: the candidate lies left of the current node's region, so the lower bound moves up; otherwise the upper bound moves down."""
    if not cur and not new:
        return 0
    if not cur:
        return 1
    if not new:
        return -1
    s1 = "".join("%03d" % i for i in cur)
    s2 = "".join("%03d" % i for i in new)
    return (s2 > s1) - (s2 < s1)


class _Grid(object):
    """llh[n, k] for data n >= first over the chain's current node list, as one device call per tree version."""

    def __init__(self, tssb):
        self.tssb = tssb
        self.refreshes = 0
        self.refresh(0)

    def refresh(self, first):
        t = self.tssb
        self.nodes = t.backend.sync_tree(t)
        self.col = {id(nd): k for k, nd in enumerate(self.nodes)}
        self.first = first
        self.llh = t.backend.llh_rows(np.arange(first, t.num_data, dtype=np.int32))
        self.spawned = getattr(t, "n_spawned", 0)
        self.refreshes += 1

    def get(self, n, node):
        return self.llh[n - self.first, self.col[id(node)]]


def resample_assignments(tssb):
    rng, root_entry = tssb.rng, tssb.root
    grid = _Grid(tssb)
    for n in range(tssb.num_data):
        datum = tssb.data[n]
        cur = tssb.assignments[n]
        # child-index path of the current node (tssb.py:101-107)
        indices, entry = [], root_entry
        for anc in cur.get_ancestors()[1:]:
            k = next(i for i, c in enumerate(entry["children"]) if c["node"] is anc)
            indices.append(k)
            entry = entry["children"][k]
        lo, hi = 0.0, 1.0
        llh_s = np.log(rng.rand()) + grid.get(n, cur)
        while True:
            u = (hi - lo) * rng.rand() + lo
            cand, path = tssb.find_node(u)
            if cand.parent() is None:                       # keep the root empty (tssb.py:118-120)
                cand = cand.children()[0]
                path = [0]
            if getattr(tssb, "n_spawned", 0) != grid.spawned:
                grid.refresh(n)                             # find_node spawned: new columns, and the parents' pi moved
            if grid.get(n, cand) > llh_s:
                if cand is not cur:
                    cur.remove_datum(n)
                    cand.add_datum(n)
                    tssb.assignments[n] = cand
                    datum.node = cand
                    if datum.index >= tssb.backend.n_ssm and n + 1 < tssb.num_data:
                        grid.refresh(n + 1)                 # a CNV moved: the rows of its SSMs still to come change
                break
            if abs(hi - lo) < EPS:
                sys.stderr.write("Slice sampler shrank down.  Keep current state.\n")
                break
            if _path_cmp(indices, path) < 0:
                lo = u
            else:
                hi = u
    tssb.last_grid_refreshes = grid.refreshes


def complete_data_log_likelihood(tssb):
    weights, nodes = tssb.get_mixture()
    tssb.backend.sync_tree(tssb)
    total = 0.0
    for w, nd in zip(weights, nodes):
        if nd.num_local_data():
            total += nd.num_local_data() * np.log(w)
    return float(total + tssb.backend.total_llh())
