"""The chain's compute backend: flattens the host tree into the arrays of the C ABI and forwards every likelihood / MH request
to the device context (phylowgs_b200.engine.Engine over libpwgs_b200.so).  This replaces the reference's file round trip
(params.write_tree / write_data_state -> mh.o -> c_params.txt, params.py:30-194).  There is no CPU implementation here."""
import numpy as np

from .. import SEM_PY
from ..engine import Engine


def arrays_from_data(data):
    """Datum list (SSMs s0.. then CNVs c0.., util2.py:53-94 order) -> the arrays of pwgs_create."""
    n_ssm = sum(1 for d in data if d.id[0] == "s")
    for i, d in enumerate(data):
        want = ("s%d" % i) if i < n_ssm else ("c%d" % (i - n_ssm))
        if d.id != want:
            raise ValueError("data must be ordered s0,s1,...,c0,c1,... (README.md:23-25); found %r at position %d" % (d.id, i))
        d.index = i
    a = np.array([d.a for d in data], dtype=np.int32)
    dd = np.array([d.d for d in data], dtype=np.int32)
    mu_r = np.array([d.mu_r for d in data], dtype=np.float64)
    mu_v = np.array([d.mu_v for d in data], dtype=np.float64)
    ptr = np.zeros(n_ssm + 1, dtype=np.int32)
    idx, c1, c2 = [], [], []
    for j in range(n_ssm):
        for cnv, x, y in data[j].cnv:
            idx.append(cnv.index - n_ssm)
            c1.append(x)
            c2.append(y)
        ptr[j + 1] = len(idx)
    return dict(a=a, d=dd, mu_r=mu_r, mu_v=mu_v, n_ssm=n_ssm, cnv_ptr=ptr, cnv_idx=np.array(idx, dtype=np.int32),
                cnv_c1=np.array(c1, dtype=np.int32), cnv_c2=np.array(c2, dtype=np.int32))


def assignment_array(tssb, nodes):
    """node_of_datum (index into `nodes`) as an int32 array, or None when a datum sits in a node that is not in `nodes`.
    Rebuilt from the nodes' data sets only when an assignment changed since the last call (Node.add_datum / remove_datum bump
    tssb._assign_version; the native sweep stores what it knows): otherwise the cached array is re-indexed to the given node
    order."""
    pos = {id(nd): k for k, nd in enumerate(nodes)}
    nod = None
    cache = tssb.__dict__.get("_nod_cache")
    if cache is not None and cache[0] == getattr(tssb, "_assign_version", 0) and len(cache[2]) == tssb.num_data:
        remap = np.array([pos.get(id(nd), -1) for nd in cache[1]], dtype=np.int32)
        cand = remap[cache[2]]
        # safety net: a node's count must be the size of its data set (any edit that bypassed the version shows up here)
        if cand.min() >= 0 and np.array_equal(np.bincount(cand, minlength=len(nodes)), [len(nd.data) for nd in nodes]):
            nod = cand
    if nod is None:
        nod = np.full(tssb.num_data, -1, dtype=np.int32)
        for k, nd in enumerate(nodes):
            if nd.data:
                nod[np.fromiter(nd.data, dtype=np.int64, count=len(nd.data))] = k
        if tssb.num_data and nod.min() < 0:
            return None
    tssb.__dict__["_nod_cache"] = (getattr(tssb, "_assign_version", 0), nodes, nod)
    return nod


def flatten_tree(tssb):
    """Preorder node list (get_mixture order = the ids evolve.py:175-177 assigns) -> (nodes, parent, node_of_datum, phi, pi).
    The tree is flattened three times per MCMC iteration; node_of_datum comes from `assignment_array`'s cache."""
    nodes = tssb.get_nodes()                               # preorder, as get_mixture lists them (without computing the weights)
    pos = {id(nd): k for k, nd in enumerate(nodes)}
    parent = np.array([-1 if nd.parent() is None else pos[id(nd.parent())] for nd in nodes], dtype=np.int32)
    nod = assignment_array(tssb, nodes)
    if nod is None:
        raise ValueError("a datum is assigned to a node outside the tree")
    phi = np.array([np.asarray(nd.params, dtype=np.float64).reshape(-1) for nd in nodes])
    pi = np.array([np.asarray(nd.pi, dtype=np.float64).reshape(-1) for nd in nodes])
    return nodes, parent, nod, phi, pi


class Backend(object):
    """Interface the host samplers use.  `GpuBackend` is the product; the parity tests plug the CPU oracle in behind the same
    interface (tests/oracle_backend.py) to obtain the chain the GPU chain must reproduce."""

    n_ssm = 0

    def set_tree_arrays(self, parent, node_of_datum, phi, pi):
        raise NotImplementedError

    def llh_rows(self, rows):
        raise NotImplementedError

    def llh_block(self, rows, cols):
        raise NotImplementedError

    def total_llh(self):
        raise NotImplementedError

    def mh(self, iters, std, seed, stream):
        raise NotImplementedError

    # column-major forms ([node][row]: one contiguous column per node, what the assignment sweep reads)
    def llh_rows_t(self, rows):
        return np.ascontiguousarray(self.llh_rows(rows).T)

    def llh_block_t(self, rows, cols):
        return np.ascontiguousarray(self.llh_block(rows, cols).T)

    # tree edits in the middle of a sweep; the defaults re-send everything through set_tree_arrays (what the CPU oracle behind
    # this interface needs), the GPU backend sends only what changed
    def grid_t(self, n_data, spare=0):
        """[K + spare][n_data]: every datum in every node (start of a sweep) in the first K rows; the spare rows are storage for
        the columns of nodes spawned during the sweep."""
        g = self.llh_rows_t(np.arange(n_data, dtype=np.int32))
        if spare:
            full = np.empty((g.shape[0] + spare, n_data))
            full[:g.shape[0]] = g
            return full
        return g

    def update_tree_arrays(self, parent, node_of_datum, phi, pi):
        """The tree grew (spawned nodes appended); node_of_datum is passed for backends that need everything again."""
        self.set_tree_arrays(parent, node_of_datum, phi, pi)

    def move_datum(self, n, k, parent, node_of_datum, phi, pi):
        """Datum n moved to node k (node_of_datum already says so); the tree itself is unchanged since the last update."""
        self.set_tree_arrays(parent, node_of_datum, phi, pi)

    def llh_range_t(self, row0, n_data, cols):
        """[len(cols)][n_data - row0]: the data still to be visited in the given nodes."""
        return self.llh_block_t(np.arange(row0, n_data, dtype=np.int32), cols)

    # -- shared glue
    def sync_tree(self, tssb):
        nodes, parent, nod, phi, pi = flatten_tree(tssb)
        self.set_tree_arrays(parent, nod, phi, pi)
        self.nodes = nodes
        return nodes

    def datum_llh(self, tssb, datum, node):
        """Single (datum, node) query behind alleles.logprob (alleles.py:52-53); the samplers use the batched grid instead."""
        nodes = self.sync_tree(tssb)
        k = next(i for i, nd in enumerate(nodes) if nd is node)
        return float(self.llh_rows(np.array([datum.index], dtype=np.int32))[0, k])


class GpuBackend(Backend):
    SCRATCH_COLS = 8
    _scratch = None

    def __init__(self, data, device=0):
        arr = arrays_from_data(data)
        self.n_ssm = arr["n_ssm"]
        self.n_cnv = len(data) - self.n_ssm
        self.ntps = arr["a"].shape[1]
        self.engine = Engine(device=device, **arr)
        # plain data through exact integer totals per (node, error-rate class, sample) in the MH kernel (K0d); PWGS_MH_COLLAPSE=0
        # keeps the per-datum streaming path
        import os
        if os.environ.get("PWGS_MH_COLLAPSE", "1") != "0" and self.engine.get_option("n_classes") > 0:
            self.engine.set_option("mh_collapse", 1)
        if os.environ.get("PWGS_MH_BIG"):                   # tuning: -1 automatic, 1 / 0 force / forbid node tables in global memory
            self.engine.set_option("mh_big", int(os.environ["PWGS_MH_BIG"]))

    @property
    def native_ctx(self):
        """pwgs_ctx* for library code that talks to the device itself (the native assignment sweep's grid refreshes)."""
        return self.engine._h

    def tree_grew(self, K):
        """The library appended nodes to the device's tree (spawns of a native sweep)."""
        self.engine.K = K

    def set_tree_arrays(self, parent, node_of_datum, phi, pi):
        self.engine.set_tree(parent, node_of_datum, phi, pi)

    def llh_rows(self, rows):
        return self.engine.llh_rows(rows, SEM_PY)

    def llh_block(self, rows, cols):
        return self.engine.llh_block(rows, cols, SEM_PY)

    def llh_rows_t(self, rows):
        return self.engine.llh_rows(rows, SEM_PY, col_major=True)     # written column-major by the kernel: no transpose on the host

    def llh_block_t(self, rows, cols):
        return self.engine.llh_block(rows, cols, SEM_PY, col_major=True)

    def grid_t(self, n_data, spare=0):
        # straight into the context's page-locked buffer: 45 MB at 50k SSMs x 110 nodes arrive at PCIe speed (pageable: 10 ms)
        K = self.engine.K
        rows = K + spare + self.SCRATCH_COLS
        # page-locking costs ~1 ms per MB: ask once for what the trees of a burn-in need (a few hundred nodes) while that stays
        # under 256 MB, so that a growing tree does not pay it again sweep after sweep
        floor_rows = min(384, (256 << 20) // (8 * max(n_data, 1)))
        out = self.engine.host_buffer((max(rows, floor_rows), n_data))[:rows]
        self.engine.llh_range(0, n_data, None, SEM_PY, col_major=True, out=out[:K])
        self._scratch = out[K + spare:].reshape(-1)         # page-locked landing area for the small blocks of the sweep
        return out[:K + spare]

    def update_tree_arrays(self, parent, node_of_datum, phi, pi):
        self.engine.set_tree(parent, None, phi, pi)

    def move_datum(self, n, k, parent, node_of_datum, phi, pi):
        self.engine.move_data(n, k)

    def llh_range_t(self, row0, n_data, cols):
        n = (n_data - row0) * len(cols)
        out = self._scratch[:n].reshape(len(cols), n_data - row0) if (self._scratch is not None and n <= len(self._scratch)) else None
        return self.engine.llh_range(row0, n_data - row0, cols, SEM_PY, col_major=True, out=out)

    def total_llh(self):
        return self.engine.total_llh(SEM_PY)

    def mh(self, iters, std, seed, stream):
        return self.engine.mh_run(iters, std, seed, stream)

    def close(self):
        self.engine.close()
