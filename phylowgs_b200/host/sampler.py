"""The two likelihood-bound TSSB call sites, restructured around batched device calls.

* `resample_assignments` (tssb.py:82-150): the slice sampler over the unit interval is inherently sequential (each datum's
  move changes node occupancy, and `find_node` may spawn nodes), but the only expensive thing it asks for is
  "log-likelihood of datum n if it sat in node k".  The reference answers that with one Python `logprob` call that re-derives
  tree metadata every time (O(N + K) each, O(N^2) per sweep, data.py:33-38).  Here the datum x node grid comes from the device
  (pwgs_llh_rows once per sweep, pwgs_llh_block for what a tree edit makes necessary, see `_Grid`), and the serial loop itself
  runs natively in the library (pwgs_resample_assignments, csrc/pwgs_host_sampler.cpp) with the tree edits mirrored back here.
  `resample_assignments_py` is the same sweep in Python; both consume one stream of uniforms in the same order and are
  bit-for-bit interchangeable (tests/test_host_chain.py), which is also what the CPU-oracle chain of the parity tests runs.
  Decisions compare fp64 numbers produced by the device against host-drawn uniforms, so given the same uniforms the assignment
  indices are those of the CPU oracle.
* `complete_data_log_likelihood` (tssb.py:404-410): sum_k n_k log w_k on the host + one device reduction for the data term.
"""
import ctypes as C
import math
from collections import deque
from itertools import repeat
import os
import sys

import numpy as np

from .model import EPS


class UniformStream(object):
    """Block-buffered uniforms from the chain's generator: the only randomness a sweep consumes (Beta(1,b) draws are taken by
    inversion, model.boundbeta1), in the order documented at pwgs_resample_assignments (include/pwgs_b200.h)."""

    def __init__(self, rng, block):
        self.rng = rng
        self.buf = rng.rand(block)
        self.pos = 0
        self.n = block

    def refill(self):
        self.buf[:] = self.rng.rand(len(self.buf))
        self.pos = 0
        self.n = len(self.buf)
        return self.n

    def next(self):
        if self.pos >= self.n:
            self.refill()
        v = self.buf[self.pos]
        self.pos += 1
        return float(v)


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
    """cols[k][n]: log-likelihood of datum n if it sat in node k, filled by device calls and kept valid through the sweep.

    Nodes get stable indices for the sweep: preorder position at the start, then spawn order (pwgs_set_tree takes any node
    order), so columns never move.  What invalidates entries that are still to be read:
      * `find_node` spawned nodes: only their columns are missing.  A spawn splits its parent's pi and leaves every existing
        node's phi unchanged, and every likelihood depends on the tree only through phi (plain data directly; CNV-affected SSMs
        through the subtree sums their piecewise-constant coefficients reduce to) -> one block call, remaining rows x new nodes;
      * a CNV datum moved: the rows of the SSMs linked to it that are still to come -> one block call, those rows x all nodes;
      * an SSM moved: nothing (no other datum's likelihood looks at it).
    """

    def __init__(self, tssb, cap_nodes=0):
        self.tssb = t = tssb
        self.backend = b = t.backend
        from .backend import flatten_tree
        nodes, parent, nod, phi, pi = flatten_tree(t)       # preorder
        b.set_tree_arrays(parent, nod, phi, pi)             # the device tree now equals the host tree
        b.nodes = nodes
        self.nodes = list(nodes)
        self.uid = {id(nd): k for k, nd in enumerate(self.nodes)}
        self.entries = []                                   # tree entry dict of every node, same indexing
        stack = [t.root]
        while stack:
            e = stack.pop()
            self.entries.append(e)
            stack.extend(reversed(e["children"]))
        assert all(e["node"] is nd for e, nd in zip(self.entries, self.nodes))
        self.nod = nod.copy()                             # the sweep edits it; flatten_tree keeps its own for the cache
        # the tree as arrays, kept in step with the spawns of the sweep (rows appended; a spawn also changes its parent's pi):
        # re-reading every node object at each of the ~100 refreshes of a sweep cost more than the refreshes themselves
        K = len(self.nodes)
        cap = max(64, 2 * K, cap_nodes)
        self._parent = np.zeros(cap, dtype=np.int32)
        self._phi = np.zeros((cap, phi.shape[1]))
        self._pi = np.zeros((cap, phi.shape[1]))
        self._parent[:K], self._phi[:K], self._pi[:K] = parent, phi, pi
        self.spawned = getattr(t, "n_spawned", 0)
        self.calls = 1
        # [K + spare][D]: the columns are views, no copies; the spare rows take the columns of nodes spawned during the sweep
        # (fresh 400 KB arrays per spawned node cost more in page faults than the device call that fills them)
        spare = os.environ.get("PWGS_GRID_SPARE")           # testing: 0 sends every spawned node's column to memory outside the buffer
        self._full = b.grid_t(t.num_data, spare=int(spare) if spare is not None else max(64, K // 2))
        self.cols = [self._full[k] for k in range(K)]
        if not hasattr(t, "_ssms_of_cnv"):                  # CNV datum index -> SSM datum indices linked to it
            m = {}
            for d in t.data:
                for cnv, _, _ in d.cnv:
                    m.setdefault(cnv.index, []).append(d.index)
            t._ssms_of_cnv = {k: np.array(sorted(v), dtype=np.int32) for k, v in m.items()}

    def _push_tree(self):
        K = len(self.nodes)
        self.backend.update_tree_arrays(self._parent[:K], self.nod, self._phi[:K], self._pi[:K])

    def register(self, entry):
        """A node spawned during the sweep gets the next index."""
        nd = entry["node"]
        k = len(self.nodes)
        self.uid[id(nd)] = k
        self.nodes.append(nd)
        self.entries.append(entry)
        if k >= len(self._parent):                          # grow
            cap = 2 * len(self._parent)
            self._parent = np.resize(self._parent, cap)
            self._phi = np.resize(self._phi, (cap, self._phi.shape[1]))
            self._pi = np.resize(self._pi, (cap, self._pi.shape[1]))
        pk = self.uid[id(nd.parent())]
        self._parent[k] = pk
        self._phi[k] = np.asarray(nd.params, dtype=np.float64).reshape(-1)
        self._pi[k] = np.asarray(nd.pi, dtype=np.float64).reshape(-1)
        self._pi[pk] = np.asarray(nd.parent().pi, dtype=np.float64).reshape(-1)   # the spawn took its pi out of the parent's

    def fetch_new_columns(self, n, first):
        """Columns of nodes[first:] for the data still to be visited (rows n..)."""
        self._push_tree()
        cols = np.arange(first, len(self.nodes), dtype=np.int32)
        block = self.backend.llh_range_t(n, self.tssb.num_data, cols)
        for i in range(len(cols)):
            k = len(self.cols)
            col = self._full[k] if k < self._full.shape[0] else np.empty(self.tssb.num_data)
            col[n:] = block[i]
            self.cols.append(col)
        self.spawned = getattr(self.tssb, "n_spawned", 0)
        self.calls += 1

    def after_spawn_py(self, n):
        first = len(self.nodes)
        known = self.uid
        stack = [self.tssb.root]
        while stack:                                        # preorder walk: spawn order within one find_node call is top-down
            e = stack.pop()
            if id(e["node"]) not in known:
                self.register(e)
            stack.extend(reversed(e["children"]))
        self.fetch_new_columns(n, first)

    def after_cnv_move(self, n, cnv_index):
        K = len(self.nodes)
        # (every spawn so far has been sent by fetch_new_columns; SSM moves change no other datum's likelihood)
        self.backend.move_datum(cnv_index, int(self.nod[cnv_index]), self._parent[:K], self.nod, self._phi[:K], self._pi[:K])
        rows = self.tssb._ssms_of_cnv.get(cnv_index)
        if rows is None:
            return
        rows = rows[rows > n]
        if len(rows) == 0:
            return
        block = self.backend.llh_block_t(rows, np.arange(K, dtype=np.int32))
        if K <= self._full.shape[0]:                        # every column is a row of the sweep's buffer: one scatter
            self._full[:K, rows] = block
        else:
            for k in range(K):
                self.cols[k][rows] = block[k]
        self.calls += 1

    def get(self, n, node):
        return self.cols[self.uid[id(node)]][n]


def _uniform_block(tssb):
    return 4 * tssb.num_data + 4096


def resample_assignments_py(tssb):
    """The sweep in Python (reference semantics tssb.py:82-150); consumes tssb.ustream like the native sweep does."""
    root_entry = tssb.root
    grid = _Grid(tssb)
    stream = tssb.ustream = UniformStream(tssb.rng, _uniform_block(tssb))
    n_ssm = tssb.backend.n_ssm
    paths = {}
    try:
        for n in range(tssb.num_data):
            datum = tssb.data[n]
            cur = tssb.assignments[n]
            indices = paths.get(id(cur))
            if indices is None:                             # child-index path of the node (tssb.py:101-107); spawns append
                indices, entry = [], root_entry             # children at the end, so paths stay valid through the sweep
                for anc in cur.get_ancestors()[1:]:
                    k = next(i for i, c in enumerate(entry["children"]) if c["node"] is anc)
                    indices.append(k)
                    entry = entry["children"][k]
                paths[id(cur)] = indices
            lo, hi = 0.0, 1.0
            llh_s = math.log(stream.next()) + grid.get(n, cur)
            while True:
                u = (hi - lo) * stream.next() + lo
                cand, path = tssb.find_node(u)
                if cand.parent() is None:                   # keep the root empty (tssb.py:118-120)
                    cand = cand.children()[0]
                    path = [0]
                if getattr(tssb, "n_spawned", 0) != grid.spawned:
                    grid.after_spawn_py(n)
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
    finally:
        tssb.ustream = None
    tssb.last_grid_refreshes = grid.calls


# ------------------------------------------------------------------ native sweep (libpwgs_b200.so, host code)
_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32)
_CB0 = C.CFUNCTYPE(C.c_int, C.c_void_p)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class _Sweep(C.Structure):
    _fields_ = [("n_data", C.c_int32), ("n_ssm", C.c_int32), ("n_nodes", C.c_int32), ("cap_nodes", C.c_int32),
                ("parent", _ip), ("main", _dp), ("child_ptr", _ip), ("child_idx", _ip), ("child_stick", _dp),
                ("assignments", _ip), ("cols", C.POINTER(_dp)), ("uniforms", _dp), ("n_uniforms", C.c_int32),
                ("dp_alpha", C.c_double), ("dp_gamma", C.c_double), ("alpha_decay", C.c_double),
                ("min_depth", C.c_int32), ("max_depth", C.c_int32), ("cap_spawn", C.c_int32), ("n_spawned", C.c_int32),
                ("spawn_parent", _ip), ("spawn_stick", _dp), ("spawn_main", _dp), ("spawn_frac", _dp),
                ("user", C.c_void_p), ("on_spawn", _CB), ("on_cnv_move", _CB), ("refill", _CB0),
                ("n_moved", C.c_int32), ("n_shrunk", C.c_int32), ("uniforms_used", C.c_int32),
                ("ctx", C.c_void_p), ("ntps", C.c_int32), ("cap_cols", C.c_int32), ("tree_parent", _ip), ("phi", _dp), ("pi", _dp),
                ("grid", _dp), ("ssm_ptr", _ip), ("ssm_idx", _ip), ("scratch", _dp), ("scratch_len", C.c_int64),
                ("n_refreshes", C.c_int32)]


def resample_assignments_native(tssb):
    from .. import _lib
    from .model import _entry
    import time
    timing = os.environ.get("PWGS_SWEEP_TIMING") == "1"
    t_start = time.perf_counter()
    t_cb = [0.0, 0, 0, 0.0]                             # seconds inside callbacks, spawn events, CNV moves, seconds mirroring spawns
    lib = _lib.load()
    cap_spawn = 8192
    # the library refreshes the grid itself after spawns / CNV moves when the backend is a device context (no callback into
    # Python per tree edit); other backends (the CPU oracle of the parity tests) are asked through the callbacks
    ctx = None if os.environ.get("PWGS_SWEEP_CALLBACKS", "0") == "1" else getattr(tssb.backend, "native_ctx", None)
    grid = _Grid(tssb, cap_nodes=(len(tssb.get_nodes()) + cap_spawn) if ctx else 0)
    t_grid = time.perf_counter()
    N, K0 = tssb.num_data, len(grid.nodes)
    cap_nodes = K0 + cap_spawn
    stream = UniformStream(tssb.rng, _uniform_block(tssb))

    parent = grid._parent[:K0].copy()
    main = np.array([e["main"] for e in grid.entries], dtype=np.float64)
    child_ptr = np.zeros(K0 + 1, dtype=np.int32)
    child_idx, child_stick = [], []
    uid = grid.uid
    for k, e in enumerate(grid.entries):
        if e["children"]:
            child_idx += [uid[id(c["node"])] for c in e["children"]]
            child_stick += e["sticks"][:len(e["children"]), 0].tolist()
        child_ptr[k + 1] = len(child_idx)
    child_idx = np.array(child_idx if child_idx else [0], dtype=np.int32)
    child_stick = np.array(child_stick if child_stick else [0.0], dtype=np.float64)
    assign = grid.nod.copy()
    before = assign.copy()
    # cols[k] = address of node k's column: rows of the sweep's buffer (one vector expression), library-owned memory beyond
    col_addr = np.zeros(cap_nodes, dtype=np.uint64)
    full_rows = grid._full.shape[0]
    col_addr[:full_rows] = grid._full.ctypes.data + np.arange(full_rows, dtype=np.uint64) * np.uint64(grid._full.strides[0])
    assert K0 <= full_rows
    col_addr[K0:] = 0
    cols = C.cast(col_addr.ctypes.data, C.POINTER(_dp))
    sp_parent = np.zeros(cap_spawn, dtype=np.int32)
    sp_stick, sp_main, sp_frac = np.zeros(cap_spawn), np.zeros(cap_spawn), np.zeros(cap_spawn)
    failure = []
    state = {"mirrored": 0}
    sw = _Sweep()

    def mirror_spawns():
        """Create the Python twins of the nodes the native sweep spawned since the last call (alleles.py:35-37, tssb.py:353-361)."""
        for i in range(state["mirrored"], sw.n_spawned):
            pe = grid.entries[int(sp_parent[i])]
            pe["sticks"] = np.vstack([pe["sticks"], float(sp_stick[i])])
            child = _entry(pe["node"].spawn(frac=float(sp_frac[i])), float(sp_main[i]))
            pe["children"].append(child)
            tssb.n_spawned = getattr(tssb, "n_spawned", 0) + 1
            grid.register(child)
        state["mirrored"] = sw.n_spawned

    def on_spawn(_user, n):
        t0 = time.perf_counter()
        try:
            first = len(grid.nodes)
            mirror_spawns()
            t_cb[3] += time.perf_counter() - t0
            grid.fetch_new_columns(int(n), first)
            for k in range(first, len(grid.nodes)):
                col_addr[k] = grid.cols[k].ctypes.data
            return 0
        except BaseException as e:          # noqa: B902 - must not propagate through the C frame
            failure.append(e)
            return -1
        finally:
            t_cb[0] += time.perf_counter() - t0
            t_cb[1] += 1

    def on_cnv_move(_user, n):
        t0 = time.perf_counter()
        try:
            grid.nod[n] = assign[n]                         # the device needs the CNV's new node (SSM moves change no likelihood)
            grid.after_cnv_move(int(n), int(n))
            return 0
        except BaseException as e:          # noqa: B902
            failure.append(e)
            return -1
        finally:
            t_cb[0] += time.perf_counter() - t0
            t_cb[2] += 1

    def refill(_user):
        try:
            return stream.refill()
        except BaseException as e:          # noqa: B902
            failure.append(e)
            return 0

    cb_spawn, cb_cnv, cb_refill = _CB(on_spawn), _CB(on_cnv_move), _CB0(refill)
    if ctx:
        cb_spawn, cb_cnv = _CB(), _CB()                    # NULL: the library does it
        if not hasattr(tssb, "_ssms_csr"):                  # CSR over the CNV data: the SSMs linked to each, ascending
            n_ssm, n_cnv = tssb.backend.n_ssm, N - tssb.backend.n_ssm
            ptr = np.zeros(n_cnv + 1, dtype=np.int32)
            chunks = []
            for j in range(n_cnv):
                rows = tssb._ssms_of_cnv.get(n_ssm + j)
                if rows is not None:
                    chunks.append(rows)
                    ptr[j + 1] = len(rows)
            np.cumsum(ptr, out=ptr)
            tssb._ssms_csr = (ptr, np.concatenate(chunks).astype(np.int32) if chunks else np.zeros(1, dtype=np.int32))
        ssm_ptr, ssm_idx = tssb._ssms_csr
        sw.ctx, sw.ntps, sw.cap_cols = ctx, grid._phi.shape[1], grid._full.shape[0]
        assert grid._parent.shape[0] >= cap_nodes and grid._full.flags.c_contiguous
        sw.tree_parent, sw.phi, sw.pi = grid._parent.ctypes.data_as(_ip), grid._phi.ctypes.data_as(_dp), grid._pi.ctypes.data_as(_dp)
        sw.grid = grid._full.ctypes.data_as(_dp)
        sw.ssm_ptr, sw.ssm_idx = ssm_ptr.ctypes.data_as(_ip), ssm_idx.ctypes.data_as(_ip)
        scratch = getattr(tssb.backend, "_scratch", None)
        if scratch is not None and len(scratch):
            sw.scratch, sw.scratch_len = scratch.ctypes.data_as(_dp), len(scratch)
    sw.n_data, sw.n_ssm, sw.n_nodes, sw.cap_nodes = N, tssb.backend.n_ssm, K0, cap_nodes
    sw.parent, sw.main = parent.ctypes.data_as(_ip), main.ctypes.data_as(_dp)
    sw.child_ptr, sw.child_idx, sw.child_stick = child_ptr.ctypes.data_as(_ip), child_idx.ctypes.data_as(_ip), child_stick.ctypes.data_as(_dp)
    sw.assignments, sw.cols = assign.ctypes.data_as(_ip), cols
    sw.uniforms, sw.n_uniforms = stream.buf.ctypes.data_as(_dp), stream.n
    sw.dp_alpha, sw.dp_gamma, sw.alpha_decay = tssb.dp_alpha, tssb.dp_gamma, tssb.alpha_decay
    sw.min_depth, sw.max_depth = tssb.min_depth, tssb.max_depth
    sw.cap_spawn, sw.n_spawned = cap_spawn, 0
    sw.spawn_parent, sw.spawn_stick = sp_parent.ctypes.data_as(_ip), sp_stick.ctypes.data_as(_dp)
    sw.spawn_main, sw.spawn_frac = sp_main.ctypes.data_as(_dp), sp_frac.ctypes.data_as(_dp)
    sw.user, sw.on_spawn, sw.on_cnv_move, sw.refill = None, cb_spawn, cb_cnv, cb_refill
    t_setup = time.perf_counter()
    rc = lib.pwgs_resample_assignments(C.cast(C.byref(sw), C.c_void_p))
    t_native = time.perf_counter()
    if failure:
        raise failure[0]
    if rc != 0:
        raise RuntimeError("pwgs_resample_assignments failed with code %d%s" % (rc, (": " + lib.pwgs_last_error().decode()) if rc == -8 else ""))
    t_m = time.perf_counter()
    mirror_spawns()
    if ctx:
        t_cb[3] += time.perf_counter() - t_m
        grid.nod[:] = assign
        grid.calls += sw.n_refreshes
        tssb.backend.tree_grew(len(grid.nodes))
    for _ in range(sw.n_shrunk):
        sys.stderr.write("Slice sampler shrank down.  Keep current state.\n")
    # mirror the moves back into the Python objects: set operations per node, two stores per moved datum
    changed = np.nonzero(assign != before)[0]
    if len(changed):
        for key, op in ((before[changed], "difference_update"), (assign[changed], "update")):
            order = np.argsort(key, kind="stable")
            ks, starts = np.unique(key[order], return_index=True)
            for k, chunk in zip(ks.tolist(), np.split(changed[order], starts[1:])):
                getattr(grid.nodes[k].data, op)(chunk.tolist())
        # two stores per moved datum, as loops inside the interpreter's C code (deque(map(...), maxlen=0) runs the map dry)
        moved = changed.tolist()
        targets = list(map(grid.nodes.__getitem__, assign[changed].tolist()))
        deque(map(tssb.assignments.__setitem__, moved, targets), maxlen=0)
        deque(map(setattr, map(tssb.data.__getitem__, moved), repeat("node"), targets), maxlen=0)
    # the assignments as an array, for the next flatten_tree (the set operations above went around Node.add_datum)
    tssb.__dict__["_nod_cache"] = (getattr(tssb, "_assign_version", 0), list(grid.nodes), assign.copy())
    tssb.last_grid_refreshes = grid.calls
    if timing:
        t_end = time.perf_counter()
        sys.stderr.write("sweep timing: grid %.1f ms, setup %.1f ms, native sweep %.1f ms (of which callbacks %.1f ms: %d spawn events, "
                         "%d CNV moves; %.1f ms of it creating the spawned nodes' Python twins), mirror-back %.1f ms; K %d -> %d\n"
                         % (1e3 * (t_grid - t_start), 1e3 * (t_setup - t_grid), 1e3 * (t_native - t_setup), 1e3 * t_cb[0], t_cb[1], t_cb[2], 1e3 * t_cb[3],
                            1e3 * (t_end - t_native), K0, len(grid.nodes)))


def resample_assignments(tssb):
    if os.environ.get("PWGS_PY_SAMPLER", "0") == "1":
        return resample_assignments_py(tssb)
    return resample_assignments_native(tssb)


def complete_data_log_likelihood(tssb):
    weights, nodes = tssb.get_mixture()
    tssb.backend.sync_tree(tssb)
    total = 0.0
    for w, nd in zip(weights, nodes):
        if nd.num_local_data():
            total += nd.num_local_data() * np.log(w)
    return float(total + tssb.backend.total_llh())
