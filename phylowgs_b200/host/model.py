"""Host-side data model of a PhyloWGS chain, Python 3.

Mirrors the objects the reference pickles into trees.zip and that its post-processing reads (SURVEY.md 8b "Output compat"):
``tssb.TSSB`` (tssb.py:9-46), ``alleles.alleles`` over ``node.Node`` (alleles.py:11-56, node.py:6-92) and ``data.Datum``
(data.py:6-22) — same class names, same attribute names, same nested-dict tree (`root = {'node','main','sticks','children'}`).
What differs is where the arithmetic happens: no likelihood is evaluated here.  Every log-likelihood the samplers need comes
from the chain's backend (the CUDA library, phylowgs_b200/host/backend.py); the branchy stick-breaking edits stay on the host,
as north_star prescribes.

Randomness: the reference draws from numpy's global legacy generator (tssb.py:5 `from numpy.random import *`); here every
sampler takes the chain's own ``numpy.random.RandomState`` (same generator family, checkpointable with get_state()).
"""
import numpy as np
from scipy.special import gammaln

EPS = np.finfo(np.float64).eps


def boundbeta(rng, a, b):
    """Beta draw squeezed off the interval ends (util.py:29-30)."""
    return (1.0 - EPS) * (rng.beta(a, b) - 0.5) + 0.5


def boundbeta1(u, b):
    """boundbeta(1, b) from one uniform by inversion (Beta(1,b): x = 1 - u^(1/b)); the form both the Python and the native
    assignment sweep use, so that they consume the same stream of uniforms (csrc/pwgs_host_sampler.cpp)."""
    return (1.0 - EPS) * ((1.0 - u ** (1.0 / b)) - 0.5) + 0.5


def betapdfln(x, a, b):
    """log Beta(x; a, b) (util.py:26-27)."""
    return gammaln(a + b) - gammaln(a) - gammaln(b) + (a - 1.0) * np.log(x) + (b - 1.0) * np.log(1.0 - x)


def sticks_to_edges(sticks):
    """Right edges of the stick-breaking intervals (util.py:11-12)."""
    return 1.0 - np.cumprod(1.0 - np.asarray(sticks, dtype=float).reshape(-1))


class Datum(object):
    """Read counts of one SSM or CNV (data.py:6-22).  `cnv` lists (cnv_datum, c1, c2) links for an SSM (util2.py:86-90).
    `index` is the datum's row in the device arrays: SSM k -> k, CNV k -> n_ssm + k (params.py:71-76)."""

    def __init__(self, name, id, a, d, mu_r=0.0, mu_v=0.0, index=-1):
        self.name = name
        self.id = id
        self.a = list(a)
        self.d = list(d)
        self.mu_r = mu_r
        self.mu_v = mu_v
        self.index = index
        self._log_bin_norm_const = [float(gammaln(dd + 1) - gammaln(aa + 1) - gammaln(dd - aa + 1)) for aa, dd in zip(self.a, self.d)]
        self.nr = 0
        self.nv = 0
        self.node = None
        self.cnv = []
        self.tssb = None


class Node(object):
    """Tree node holding a set of datum indices (node.py:6-92)."""

    def __init__(self, parent=None, tssb=None):
        self.data = set()
        self._children = []
        self.tssb = tssb
        self._parent = parent
        if parent is not None:
            parent._children.append(self)

    # -- structure
    def parent(self):
        return self._parent

    def children(self):
        return self._children

    def add_child(self, child):
        self._children.append(child)

    def remove_child(self, child):
        self._children.remove(child)

    def get_ancestors(self):
        """Root ... self (node.py:86-92)."""
        chain, n = [], self
        while n is not None:
            chain.append(n)
            n = n._parent
        return chain[::-1]

    # -- data bookkeeping
    def add_datum(self, n):
        self.data.add(n)
        t = self.tssb
        if t is not None:                                   # invalidates the cached assignment array of backend.flatten_tree
            t._assign_version = getattr(t, "_assign_version", 0) + 1

    def remove_datum(self, n):
        self.data.remove(n)
        t = self.tssb
        if t is not None:
            t._assign_version = getattr(t, "_assign_version", 0) + 1

    def num_local_data(self):
        return len(self.data)

    def num_data(self):
        return len(self.data) + sum(c.num_data() for c in self._children)

    def has_data(self):
        return bool(self.data) or any(c.has_data() for c in self._children)

    def get_data(self):
        return [self.tssb.data[n] for n in self.data]

    def spawn(self, **kw):
        return self.__class__(parent=self, tssb=self.tssb, **kw)

    def kill(self):
        if self._parent is not None:
            self._parent._children.remove(self)
        self._parent = None
        self._children = None


class alleles(Node):
    """Subclone: population frequency `params` (phi) and own mass `pi` per sample (alleles.py:11-56).  A spawned child takes a
    uniform fraction of its parent's pi, one fraction for all samples (alleles.py:35-37); killing returns it (45-50)."""

    init_mean = 0.5
    min_conc = 0.01
    max_conc = 0.1

    def __init__(self, parent=None, tssb=None, conc=0.1, ntps=5, frac=None):
        super(alleles, self).__init__(parent=parent, tssb=tssb)
        if tssb is not None:
            ntps = len(tssb.data[0].a)
        self.params1 = np.zeros(ntps)
        self.pi1 = np.zeros(ntps)
        self.path = None
        self.ht = 0
        if parent is None:
            self._conc = conc
            self.pi = np.ones(ntps)
            self.params = np.ones(ntps)
        else:
            if frac is None:                                # the share of the parent's pi: one uniform for all samples
                stream = getattr(tssb, "ustream", None)
                frac = stream.next() if stream is not None else (tssb.rng.rand(1) if tssb is not None else np.random.rand(1))
            self.pi = frac * parent.pi
            parent.pi = parent.pi - self.pi
            self.params = self.pi

    def conc(self):
        n = self
        while n._parent is not None:
            n = n._parent
        return n._conc

    def kill(self):
        if self._parent is not None:
            self._parent._children.remove(self)
            self._parent.pi = self._parent.pi + self.pi
        self._parent = None
        self._children = None

    # alleles.py:52-56: the likelihood hooks route to the chain's backend (no arithmetic on the host)
    def logprob(self, x):
        return self.tssb.backend.datum_llh(self.tssb, x[0], self)

    def complete_logprob(self):
        return sum(self.logprob([dat]) for dat in self.get_data())

    def data_log_likelihood(self):
        return self.complete_logprob()


def _entry(node, main):
    return {"node": node, "main": main, "sticks": np.empty((0, 1)), "children": []}


class TSSB(object):
    """Tree-structured stick-breaking process over the data (tssb.py:9-46) with its host-side samplers."""

    min_dp_alpha, max_dp_alpha = 1.0, 50.0
    min_dp_gamma, max_dp_gamma = 1.0, 10.0
    min_alpha_decay, max_alpha_decay = 0.05, 0.80

    def __init__(self, dp_alpha=1.0, dp_gamma=1.0, root_node=None, data=None, min_depth=0, max_depth=15, alpha_decay=1.0,
                 rng=None, backend=None):
        if root_node is None:
            raise Exception("Root node must be specified.")
        self.rng = rng if rng is not None else np.random.RandomState()
        self.ustream = None                                 # set during an assignment sweep: the sweep's stream of uniforms
        self.backend = backend
        self.min_depth = min_depth
        self.max_depth = max_depth
        self.dp_alpha = dp_alpha
        self.dp_gamma = dp_gamma
        self.alpha_decay = alpha_decay
        self.data = data
        self.num_data = 0 if data is None else len(data)
        self.root = _entry(root_node, boundbeta(self.rng, 1.0, dp_alpha) if min_depth == 0 else 0.0)
        root_node.tssb = self
        self.assignments = []
        for n in range(self.num_data):
            root_node.add_datum(n)
            self.assignments.append(root_node)

    # the generator and the device context are process-local: keep them out of trees.zip / state pickles
    def __getstate__(self):
        st = dict(self.__dict__)
        for key in ("backend", "rng", "ustream", "_ssms_of_cnv", "_ssms_csr", "_tree_pickler", "_nod_cache", "_assign_version"):
            st.pop(key, None)
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self.backend = None
        self.rng = None
        self.ustream = None

    # ------------------------------------------------------------------ traversal
    def _new_child(self, entry, depth):
        """Break a new stick under `entry` and spawn its node (tssb.py:353-361 / 220-225)."""
        self.n_spawned = getattr(self, "n_spawned", 0) + 1
        us = getattr(self, "ustream", None)
        a_child = (float(self.alpha_decay) ** (depth + 1)) * float(self.dp_alpha)
        if us is not None:                                  # inside a sweep: stick, share of pi, main, each from one uniform
            stick = boundbeta1(us.next(), float(self.dp_gamma)) if depth != 0 else 0.999
            node = entry["node"].spawn()
            main = boundbeta1(us.next(), a_child) if self.min_depth <= depth + 1 else 0.0
        else:
            stick = boundbeta(self.rng, 1, self.dp_gamma) if depth != 0 else 0.999
            node = entry["node"].spawn()
            main = boundbeta(self.rng, 1.0, a_child) if self.min_depth <= depth + 1 else 0.0
        entry["sticks"] = np.vstack([entry["sticks"], stick])
        entry["children"].append(_entry(node, main))

    def find_node(self, u):
        """Node owning point u of the unit interval, spawning sticks on the way if u falls beyond them (tssb.py:340-377).
        Returns (node, path of child indices)."""
        entry, depth, path = self.root, 0, []
        while True:
            if depth >= self.max_depth or u < entry["main"]:
                return entry["node"], path
            u = (u - entry["main"]) / (1.0 - entry["main"])
            if depth > 0:
                # right edges of the children's intervals as a running product (plain floats: this is the hot host loop)
                rem, edges = 1.0, []
                for s in (entry["sticks"][:, 0].tolist() if len(entry["sticks"]) else ()):
                    rem *= 1.0 - s
                    edges.append(1.0 - rem)
                while not edges or edges[-1] < u:
                    self._new_child(entry, depth)
                    rem *= 1.0 - float(entry["sticks"][-1, 0])
                    edges.append(1.0 - rem)
                index = 0
                while edges[index] < u:
                    index += 1
                left = edges[index - 1] if index else 0.0
                u = (u - left) / (edges[index] - left)
            else:
                index = 0                                   # the root always descends into its first child (tssb.py:369-371)
            path.append(index)
            entry = entry["children"][index]
            depth += 1

    def get_nodes(self):
        out, stack = [], [self.root]
        while stack:
            e = stack.pop()
            out.append(e["node"])
            stack.extend(reversed(e["children"]))
        return out

    def get_mixture(self):
        """(weights, nodes) in preorder (tssb.py:389-402)."""
        weights, nodes = [], []

        def walk(entry, mass):
            weights.append(mass * entry["main"])
            nodes.append(entry["node"])
            if entry["children"]:
                # widths of the children's intervals, diff of 1 - cumprod(1 - sticks) (util.py:11-12), in plain floats: the same
                # operations in the same order as the array expression (checked in tests/test_stick_orders.py), without
                # numpy's per-call overhead on the mostly tiny stick arrays -- the tree is walked several times per MCMC iteration
                rest, rem, left = mass * (1.0 - entry["main"]), 1.0, 0.0
                for child, s in zip(entry["children"], entry["sticks"][:, 0].tolist()):
                    rem *= 1.0 - s
                    edge = 1.0 - rem
                    walk(child, rest * (edge - left))
                    left = edge
        walk(self.root, 1.0)
        return weights, nodes

    # ------------------------------------------------------------------ samplers that stay on the host (O(K))
    def cull_tree(self):
        """Drop trailing children without data anywhere below them (tssb.py:152-166)."""
        def walk(entry):
            counts = [walk(c) for c in entry["children"]]
            keep = len(counts)
            while keep > 0 and counts[keep - 1] == 0:
                keep -= 1
            for child in entry["children"][keep:]:
                child["node"].kill()
                del child["node"]
            entry["sticks"] = entry["sticks"][:keep]
            entry["children"] = entry["children"][:keep]
            return sum(counts) + entry["node"].num_local_data()
        walk(self.root)

    def resample_sticks(self):
        """Conjugate Beta updates of every stick given the data counts below it (tssb.py:168-191)."""
        def walk(entry, depth):
            below = 0
            for i in range(len(entry["children"]) - 1, -1, -1):
                n_child = walk(entry["children"][i], depth + 1)
                entry["sticks"][i] = boundbeta(self.rng, 1.0 + n_child, self.dp_gamma + below) if depth != 0 else .999999
                below += n_child
            here = entry["node"].num_local_data()
            entry["main"] = boundbeta(self.rng, 1.0 + here, (self.alpha_decay ** depth) * self.dp_alpha + below) \
                if self.min_depth <= depth else 0.0
            if depth == 0:
                entry["main"] = 1e-30                       # the root stays empty (tssb.py:187)
            return here + below
        walk(self.root, 0)

    def resample_stick_orders(self):
        """Size-biased re-ordering of the children of every node; unrepresented children are dropped (tssb.py:193-243)."""
        def walk(entry, depth):
            if not entry["children"]:
                return
            order = []
            todo = set(i for i, c in enumerate(entry["children"]) if c["node"].has_data())
            weights = None
            while todo:
                u = self.rng.rand()
                while True:
                    n_sticks = entry["sticks"].shape[0]
                    if n_sticks <= 6:
                        # plain floats for the usual handful of children: the same IEEE operations in the same order as the
                        # array code below (numpy adds fewer than 8 numbers one after the other; cumprod / cumsum always)
                        wl, prod, prev, total = [], 1.0, 0.0, 0.0
                        for st in entry["sticks"].reshape(-1).tolist():
                            prod = (1.0 - st) if not wl else prod * (1.0 - st)
                            edge = 1.0 - prod
                            wl.append(edge - prev)
                            total += wl[-1]
                            prev = edge
                        rest = [i for i in range(n_sticks) if i not in order]
                        w = [wl[i] for i in rest] + [1.0 - total]
                        wsum = 0.0
                        for x in w:
                            wsum += x
                        acc, pick = 0.0, 0
                        if wsum != 0.0:                     # (0/0 in the array code: every comparison false, pick 0)
                            for x in w:
                                acc += x / wsum
                                pick += u > acc
                    else:
                        if weights is None or len(weights) != n_sticks:
                            weights = np.diff(np.hstack([0.0, sticks_to_edges(entry["sticks"])]))
                        rest = [i for i in range(n_sticks) if i not in order]
                        w = np.hstack([weights[rest], 1.0 - np.sum(weights)])
                        w = w / np.sum(w)
                        pick = int(np.sum(u > np.cumsum(w)))
                    if pick == len(rest):                   # landed beyond the existing sticks: break a new one and retry
                        entry["sticks"] = np.vstack([entry["sticks"], boundbeta(self.rng, 1, self.dp_gamma)])
                        self.n_spawned = getattr(self, "n_spawned", 0) + 1
                        node = entry["node"].spawn()
                        main = boundbeta(self.rng, 1.0, (self.alpha_decay ** (depth + 1)) * self.dp_alpha) \
                            if self.min_depth <= depth + 1 else 0.0
                        entry["children"].append(_entry(node, main))
                        weights = None
                    else:
                        pick = rest[pick]
                        break
                order.append(pick)
                todo.discard(pick)
            kept = []
            for k in order:
                kept.append(entry["children"][k])
                walk(entry["children"][k], depth + 1)
            for k in range(entry["sticks"].shape[0]):
                if k not in order:
                    entry["children"][k]["node"].kill()
                    del entry["children"][k]["node"]
            entry["children"] = kept
            entry["sticks"] = np.zeros((len(kept), 1))
        walk(self.root, 0)
        self.resample_sticks()

    def dp_alpha_llh(self, dp_alpha, alpha_decay):
        total, stack = 0.0, [(self.root, 0)]
        while stack:
            entry, depth = stack.pop()
            if self.min_depth <= depth:
                total += betapdfln(entry["main"], 1.0, (alpha_decay ** depth) * dp_alpha)
            stack.extend((c, depth + 1) for c in entry["children"])
        return float(total)

    def dp_gamma_llh(self, dp_gamma):
        total, stack = 0.0, [self.root]
        while stack:
            entry = stack.pop()
            for i, c in enumerate(entry["children"]):
                total += float(betapdfln(float(np.ravel(entry["sticks"][i])[0]), 1.0, dp_gamma))
                stack.append(c)
        return total

    def _slice_1d(self, llh, current, lower, upper):
        """Shrinking-interval slice sampler on one hyper-parameter (tssb.py:258-272 and its two siblings)."""
        llh_s = np.log(self.rng.rand()) + llh(current)
        while True:
            cand = (upper - lower) * self.rng.rand() + lower
            if llh(cand) > llh_s:
                return cand
            if cand < current:
                lower = cand
            elif cand > current:
                upper = cand
            else:
                raise Exception("Slice sampler shrank to zero!")

    def resample_hypers(self, dp_alpha=True, alpha_decay=True, dp_gamma=True):
        """tssb.py:245-316.  The sticks do not change while the hyper-parameters are slice-sampled, so they are gathered once
        and every likelihood evaluation is one vector expression (the reference walks the tree per evaluation)."""
        mains, depths, sticks, stack = [], [], [], [(self.root, 0)]
        while stack:
            entry, depth = stack.pop()
            if self.min_depth <= depth:
                mains.append(float(entry["main"]))
                depths.append(depth)
            for i, c in enumerate(entry["children"]):
                sticks.append(float(np.ravel(entry["sticks"][i])[0]))
                stack.append((c, depth + 1))
        log1m_main = np.log(1.0 - np.array(mains))
        depths = np.array(depths, dtype=float)
        log1m_stick = np.log(1.0 - np.array(sticks)) if sticks else np.zeros(0)

        def beta1_llh(log1m, b):
            # sum of log Beta(x; 1, b) = gammaln(1+b) - gammaln(b) + (b-1) log(1-x)   (betapdfln with a = 1)
            return float(np.sum(gammaln(1.0 + b) - gammaln(b) + (b - 1.0) * log1m))

        def alpha_llh(a, decay):
            return beta1_llh(log1m_main, (decay ** depths) * a)

        if dp_alpha:
            self.dp_alpha = self._slice_1d(lambda v: alpha_llh(v, self.alpha_decay), self.dp_alpha, self.min_dp_alpha, self.max_dp_alpha)
        if alpha_decay:
            self.alpha_decay = self._slice_1d(lambda v: alpha_llh(self.dp_alpha, v), self.alpha_decay,
                                              self.min_alpha_decay, self.max_alpha_decay)
        if dp_gamma:
            self.dp_gamma = self._slice_1d(lambda v: beta1_llh(log1m_stick, np.float64(v)), self.dp_gamma, self.min_dp_gamma, self.max_dp_gamma)

    # ------------------------------------------------------------------ the two likelihood-bound call sites (device)
    def resample_assignments(self):
        from .sampler import resample_assignments
        return resample_assignments(self)

    def complete_data_log_likelihood(self):
        from .sampler import complete_data_log_likelihood
        return complete_data_log_likelihood(self)


# Names under which the reference's unpickler looks these classes up (tssb.TSSB, alleles.alleles, node.Node, data.Datum).
PICKLE_MODULES = {"TSSB": "tssb", "alleles": "alleles", "Node": "node", "Datum": "data"}
