"""resample_stick_orders: the scalar fast path for nodes with a handful of children must make the very same draws and edits
as the array formulation it replaces (kept here as the reference), stick for stick and bit for bit."""
import copy

import numpy as np

from phylowgs_b200.host.model import TSSB, Datum, _entry, alleles, boundbeta, sticks_to_edges


def reference_stick_orders(self):
    """The array formulation (tssb.py:193-243 restated), as it was before the fast path."""
    def walk(entry, depth):
        if not entry["children"]:
            return
        order = []
        todo = set(i for i, c in enumerate(entry["children"]) if c["node"].has_data())
        weights = np.diff(np.hstack([0.0, sticks_to_edges(entry["sticks"])]))
        while todo:
            u = self.rng.rand()
            while True:
                rest = [i for i in range(entry["sticks"].shape[0]) if i not in order]
                w = np.hstack([weights[rest], 1.0 - np.sum(weights)])
                w = w / np.sum(w)
                pick = int(np.sum(u > np.cumsum(w)))
                if pick == len(rest):
                    entry["sticks"] = np.vstack([entry["sticks"], boundbeta(self.rng, 1, self.dp_gamma)])
                    self.n_spawned = getattr(self, "n_spawned", 0) + 1
                    node = entry["node"].spawn()
                    main = boundbeta(self.rng, 1.0, (self.alpha_decay ** (depth + 1)) * self.dp_alpha) if self.min_depth <= depth + 1 else 0.0
                    entry["children"].append(_entry(node, main))
                    weights = np.diff(np.hstack([0.0, sticks_to_edges(entry["sticks"])]))
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


def random_tree(seed, n_nodes, n_data, fanout):
    rng = np.random.RandomState(seed)
    data = [Datum("g", "s%d" % i, [5], [10], 0.999, 0.5, i) for i in range(n_data)]
    root = alleles(conc=0.1, ntps=1)
    t = TSSB(dp_alpha=25.0, dp_gamma=1.0, alpha_decay=0.25, root_node=root, data=data, rng=rng, backend=None)
    entries = [t.root]
    for k in range(n_nodes):
        pe = entries[rng.randint(0, min(len(entries), fanout))] if fanout else entries[rng.randint(0, len(entries))]
        e = _entry(pe["node"].spawn(), float(rng.uniform(0.05, 0.9)))
        pe["children"].append(e)
        pe["sticks"] = np.vstack([pe["sticks"], rng.uniform(0.05, 0.9)])
        entries.append(e)
    holders = entries[1:]
    for n in range(n_data):
        nd = holders[rng.randint(0, len(holders)) if n % 3 else rng.randint(0, max(1, len(holders) // 2))]["node"]   # some nodes stay empty
        root.remove_datum(n)
        nd.add_datum(n)
        t.assignments[n] = nd
    return t


def shape(t):
    out, stack = [], [(t.root, 0)]
    while stack:
        e, depth = stack.pop()
        out.append((depth, sorted(e["node"].data), float(e["main"]), e["sticks"].reshape(-1).tolist(), len(e["children"])))
        stack.extend((c, depth + 1) for c in reversed(e["children"]))
    return out


def test_fast_path_equals_the_array_formulation():
    for seed in range(40):
        n_nodes = [3, 8, 25, 60][seed % 4]
        fanout = [0, 2, 12, 1][(seed // 4) % 4]                    # bushy, binary-ish, one node with many children, chain
        a = random_tree(seed, n_nodes, 90, fanout)
        b = copy.deepcopy(a)                                       # (the tree's pickling hooks leave the generator behind)
        b.rng = np.random.RandomState()
        b.rng.set_state(a.rng.get_state())
        a.resample_stick_orders()
        reference_stick_orders(b)
        assert shape(a) == shape(b), seed
        assert a.rng.rand() == b.rng.rand(), seed                  # the generators consumed the same draws
        assert getattr(a, "n_spawned", 0) == getattr(b, "n_spawned", 0)


def test_get_mixture_in_plain_floats_equals_the_array_expression():
    """get_mixture's scalar walk against tssb.py:389-402 written with arrays (diff of 1 - cumprod(1 - sticks)): same bits."""
    t = random_tree(8, 60, 30, 0)
    want_w, want_n = [], []

    def walk(entry, mass):
        want_w.append(mass * entry["main"])
        want_n.append(entry["node"])
        w = np.diff(np.hstack([0.0, sticks_to_edges(entry["sticks"])]))
        for i, child in enumerate(entry["children"]):
            walk(child, mass * (1.0 - entry["main"]) * w[i])
    walk(t.root, 1.0)
    got_w, got_n = t.get_mixture()
    assert all(a is b for a, b in zip(got_n, want_n)) and len(got_n) == 61
    assert [float(x).hex() for x in got_w] == [float(x).hex() for x in want_w]
    assert got_n == t.get_nodes()
