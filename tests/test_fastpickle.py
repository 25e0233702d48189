"""The hand-assembled protocol-2 pickle of a TSSB (host/fastpickle.py) against the ordinary pickler: same object graph after
`pickle.loads`, with plain `pickle` on the reading side (what util2.TreeReader does)."""
import os
import pickle
import pickletools

import numpy as np
import pytest

from phylowgs_b200.host import evolve as E
from phylowgs_b200.host import fastpickle
from phylowgs_b200.host.model import TSSB, alleles, boundbeta
from tests import cases

SSM = os.path.join(cases.GOLDEN, "ssm_data.txt")
CNV = os.path.join(cases.GOLDEN, "cnv_data.txt")


def make_tree(ssm=SSM, cnv=CNV, n_children=3, seed=4):
    E.install_pickle_aliases()
    codes, n_ssms, n_cnvs, _ = E.load_data(ssm, cnv)
    rng = np.random.RandomState(seed)
    root = alleles(conc=0.1, ntps=len(codes[0].a))
    t = TSSB(dp_alpha=25.0, dp_gamma=1.0, alpha_decay=0.25, root_node=root, data=codes, rng=rng, backend=None)
    kids = []
    for k in range(n_children):
        parent = root if k < 2 else kids[0]["node"]
        entry = {"node": parent.spawn(), "main": boundbeta(rng, 1.0, 2.0), "sticks": np.empty((0, 1)), "children": []}
        (t.root if k < 2 else kids[0])["children"].append(entry)
        holder = t.root if k < 2 else kids[0]
        holder["sticks"] = np.vstack([holder["sticks"], rng.rand()])
        kids.append(entry)
    for n in range(t.num_data):
        nd = kids[n % n_children]["node"]
        root.remove_datum(n)
        nd.add_datum(n)
        t.assignments[n] = nd
    for d in codes:
        d.tssb = t
    E.map_datum_to_node(t)
    return t


def same_graph(a, b):
    assert type(a) is type(b) and a.num_data == b.num_data
    assert set(a.__dict__) - {"_tree_pickler"} == set(b.__dict__) - {"_tree_pickler"}
    na, nb = a.get_nodes(), b.get_nodes()
    assert len(na) == len(nb)
    ida, idb = {id(n): i for i, n in enumerate(na)}, {id(n): i for i, n in enumerate(nb)}
    for x, y in zip(na, nb):
        assert x.data == y.data and x.tssb is a and y.tssb is b
        assert np.array_equal(x.params, y.params) and np.array_equal(x.pi, y.pi)
    assert [ida[id(n)] for n in a.assignments] == [idb[id(n)] for n in b.assignments]
    assert len(a.data) == len(b.data)
    pa, pb = {id(d): i for i, d in enumerate(a.data)}, {id(d): i for i, d in enumerate(b.data)}
    for da, db in zip(a.data, b.data):
        assert type(da) is type(db) and da.tssb is a and db.tssb is b
        assert set(da.__dict__) == set(db.__dict__)
        for key, va in da.__dict__.items():
            vb = db.__dict__[key]
            if key == "tssb":
                continue
            if key == "node":
                assert (va is None) == (vb is None) and (va is None or ida[id(va)] == idb[id(vb)])
                assert va is None or va is a.assignments[pa[id(da)]]          # the very object, not a copy
            elif key == "cnv":
                assert [(pa[id(c)], x, y) for c, x, y in va] == [(pb[id(c)], x, y) for c, x, y in vb]
                assert all(isinstance(l, tuple) for l in va)
            else:
                assert type(va) is type(vb) and va == vb, key


def test_fast_pickle_unpickles_to_the_same_tree():
    t = make_tree()
    plain = pickle.loads(pickle.dumps(t, protocol=2))
    blob = fastpickle.dumps(t)
    assert t.__dict__["_tree_pickler"].ok                                   # the fast path ran (no fallback)
    pickletools.dis(blob, out=open(os.devnull, "w"))                       # a well-formed stream for the generic reader
    assert blob[:2] == b"\x80\x02"
    fast = pickle.loads(blob)
    same_graph(fast, plain)
    assert isinstance(t.data, list) and t.data[0].tssb is t                 # the live tree is untouched
    assert "_tree_pickler" not in fast.__dict__


def test_fast_pickle_follows_the_chain_and_only_uses_protocol_2_opcodes():
    t = make_tree(n_children=4, seed=9)
    nodes = t.get_nodes()
    for step in range(3):                                                   # move data around, spawn a node, pickle again
        for n in range(step, t.num_data, 3):
            old, new = t.assignments[n], nodes[1 + (n + step) % (len(nodes) - 1)]
            old.remove_datum(n)
            new.add_datum(n)
            t.assignments[n] = new
        E.map_datum_to_node(t)
        fast = pickle.loads(fastpickle.dumps(t))
        same_graph(fast, pickle.loads(pickle.dumps(t, protocol=2)))
    ops = {op.name for op, _, _ in pickletools.genops(fastpickle.dumps(t))}
    proto2 = {o.name for o in pickletools.opcodes if o.proto <= 2}
    assert ops <= proto2, ops - proto2


def test_native_join_writes_the_bytes_of_the_python_join():
    """pwgs_pickle_join (host code of the library) against the Python join of the same fragments, through tree edits."""
    t = make_tree(n_children=4, seed=5)
    nodes = t.get_nodes()
    E.map_datum_to_node(t)
    for step in range(3):
        for n in range(step, t.num_data, 2):
            old, new = t.assignments[n], nodes[1 + (n + 2 * step) % (len(nodes) - 1)]
            old.remove_datum(n)
            new.add_datum(n)
            t.assignments[n] = new
            t.data[n].node = new
        native = fastpickle.dumps(t)
        tp = t.__dict__["_tree_pickler"]
        assert tp.ok and isinstance(native, bytearray)                      # the native join ran
        tp._join_native = lambda *a: None
        try:
            python = fastpickle.dumps(t)
        finally:
            del tp._join_native
        assert isinstance(python, bytes) and python == native
    t.data[0].node = nodes[0]                                               # datum.node out of step with the assignments:
    blob = fastpickle.dumps(t)
    assert isinstance(blob, bytes)                                          # the shortcut steps aside
    got = pickle.loads(blob)
    assert got.data[0].node is got.root["node"] and got.assignments[0] is not got.root["node"]


def test_fast_pickle_without_cnvs_and_with_unassigned_nodes(tmp_path):
    empty = str(tmp_path / "empty_cnv.txt")
    open(empty, "w").close()
    t = make_tree(ssm=os.path.join(cases.GOLDEN, "emptysim.3.20.10.ssm.txt"), cnv=empty)
    for d in t.data[::2]:
        d.node = None                                                       # before the first map_datum_to_node
    same_graph(pickle.loads(fastpickle.dumps(t)), pickle.loads(pickle.dumps(t, protocol=2)))


def test_fallback_when_a_datum_is_not_plain():
    t = make_tree()
    t.data[3].extra = np.arange(3)                                          # not a plain value: the ordinary pickler takes over
    blob = fastpickle.dumps(t)
    assert not t.__dict__["_tree_pickler"].ok
    assert np.array_equal(pickle.loads(blob).data[3].extra, np.arange(3))


def test_state_dict_holding_the_tree():
    """state.last.pickle: a dict with the tree inside (evolve.py StateManager) goes through the same splice."""
    t = make_tree()
    state = {"tssb": t, "last_iteration": 7, "mh_std": 200.0, "trace": np.arange(5.0), "glist": [d.name for d in t.data]}
    got = pickle.loads(fastpickle.dumps(t, root=state))
    want = pickle.loads(pickle.dumps(state, protocol=2))
    assert set(got) == set(want) and got["last_iteration"] == 7 and np.array_equal(got["trace"], want["trace"]) and got["glist"] == want["glist"]
    same_graph(got["tssb"], want["tssb"])


def test_fragment_building_is_thread_safe():
    """Chains of one process (multievolve --in-process) build their caches at the same time."""
    import threading
    trees = [make_tree(ssm=os.path.join(cases.GOLDEN, "emptysim.4.100.50.ssm.txt"), cnv=os.devnull, seed=s) for s in range(4)]
    want = [pickle.dumps(t, protocol=2) for t in trees]
    blobs, errors = [None] * len(trees), []

    def work(i):
        try:
            for _ in range(3):
                trees[i].__dict__.pop("_tree_pickler", None)          # rebuild the cache every time
                blobs[i] = fastpickle.dumps(trees[i])
        except Exception as e:                                         # noqa: BLE001
            errors.append(repr(e))
    import sys
    old = sys.getswitchinterval()
    sys.setswitchinterval(1e-6)                                        # hand the GIL over as often as possible
    try:
        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(trees))]
        for th in threads:
            th.start()
        for th in threads:
            th.join(120)
    finally:
        sys.setswitchinterval(old)
    assert not errors, errors
    for t, blob, w in zip(trees, blobs, want):
        assert t.__dict__["_tree_pickler"].ok
        same_graph(pickle.loads(blob), pickle.loads(w))


def test_pure_python_unpickler_reads_it_too():
    """The pure-Python unpickler (dict memo, opcode by opcode -- the closest stand-in here for Python 2's pickle module)."""
    import io
    t = make_tree(n_children=4, seed=12)
    got = pickle._Unpickler(io.BytesIO(fastpickle.dumps(t))).load()
    same_graph(got, pickle.loads(pickle.dumps(t, protocol=2)))
