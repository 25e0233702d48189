"""post_assign_ssm: the trees.zip reader with empty-vertex removal and the per-tree slice-sampling step, with the
CPU oracle behind the Backend interface; the -m gpu test demands the same assignments from the device."""
import json
import os
import zipfile

import numpy as np
import pytest

from phylowgs_b200.host import fastpickle
from phylowgs_b200.host import post_assign_ssm as PA
from tests import cases
from tests.oracle_backend import OracleBackend
from tests.test_fastpickle import CNV, SSM, make_tree


def write_zip(path, trees):
    with zipfile.ZipFile(path, "w", compression=zipfile.ZIP_DEFLATED) as z:
        for i, t in enumerate(trees):
            idx = i - 2                                       # two burn-in trees, then samples 0..
            z.writestr("%s_%d_%s" % ("burnin" if idx < 0 else "tree", idx, -1000.0 - i), fastpickle.dumps(t))
        z.writestr("params.json", "{}")
    return path


@pytest.fixture()
def run_zip(tmp_path):
    trees = [make_tree(n_children=3 + (s % 2), seed=20 + s) for s in range(5)]
    return write_zip(str(tmp_path / "trees.zip"), trees), trees


def test_reader_orders_trees_and_removes_empty_vertices(tmp_path):
    t = make_tree(n_children=4, seed=2)
    nodes = t.get_nodes()
    victim = nodes[1]                                         # empty this inner node: its child must move up to the root
    target = nodes[-1]
    for n in list(victim.data):
        victim.remove_datum(n)
        target.add_datum(n)
        t.assignments[n] = target
    from phylowgs_b200.host import evolve as E
    E.map_datum_to_node(t)
    n_before = len(t.get_nodes())
    path = write_zip(str(tmp_path / "z.zip"), [t, t, t])
    got = list(PA.read_trees(path))
    assert [i for i, _ in got] == [-2, -1, 0]
    for _, tree in got:
        kept = tree.get_nodes()
        assert len(kept) == n_before - 1 and all(nd.data or nd.parent() is None for nd in kept)
        assert sum(len(nd.data) for nd in kept) == tree.num_data
        stack = [tree.root]
        while stack:                                          # sticks and children stay in step, parents are consistent
            e = stack.pop()
            assert len(e["sticks"]) == len(e["children"])
            for c in e["children"]:
                assert c["node"].parent() is e["node"]
                stack.append(c)


def test_assignments_with_the_oracle_backend(run_zip):
    path, trees = run_zip
    ssms = PA.read_requested_ssms(SSM, CNV, ["s2", "s7", "s4"])
    assert [s[1] for s in ssms] == ["s2", "s4", "s7"] and ssms[0][6] == [("c0", 1, 2)] and ssms[2][6] == []
    a = PA.post_assignments(ssms, path, OracleBackend, np.random.RandomState(5))
    b = PA.post_assignments(ssms, path, OracleBackend, np.random.RandomState(5))
    assert a == b and set(a) == {"s2", "s4", "s7"}
    for sid, where in a.items():
        assert len(where) == len(trees)
        assert all(0 < w < 6 for w in where)                  # never the (empty) root; index in the descending-phi walk
    c = PA.post_assignments(ssms, path, OracleBackend, np.random.RandomState(6))
    assert set(c) == set(a)


def test_device_data_layout():
    t = make_tree()
    ssms = PA.read_requested_ssms(SSM, CNV, ["s2", "s7"])
    data, perm, new_rows = PA.device_data(t.data, ssms)
    n_ssm = sum(1 for d in t.data if d.id[0] == "s")
    assert [d.id for d in data] == ["s%d" % i for i in range(n_ssm + 2)] + ["c%d" % i for i in range(len(t.data) - n_ssm)]
    assert list(new_rows) == [n_ssm, n_ssm + 1]
    for i, d in enumerate(t.data):                            # the old data kept their numbers and their links, under the new rows
        y = data[perm[i]]
        assert (y.a, y.d, y.mu_r, y.mu_v) == (d.a, d.d, d.mu_r, d.mu_v)
        assert [(c.name, c1, c2) for c, c1, c2 in y.cnv] == [(c.name, c1, c2) for c, c1, c2 in d.cnv]
    assert [(c.id, c1, c2) for c, c1, c2 in data[n_ssm].cnv] == [("c0", 1, 2)] and data[n_ssm + 1].cnv == []


@pytest.mark.gpu
def test_device_gives_the_oracle_assignments(run_zip, gpu_lib, capsys):
    path, trees = run_zip
    ssms = PA.read_requested_ssms(SSM, CNV, ["s2", "s7", "s4"])
    want = PA.post_assignments(ssms, path, OracleBackend, np.random.RandomState(11))
    got = PA.post_assignments(ssms, path, PA.gpu_backend_factory(0), np.random.RandomState(11))
    assert got == want
    PA.main(["--cnvs", CNV, "--seed", "11", SSM, path, "s2", "s7", "s4"])
    out = json.loads(capsys.readouterr().out)
    assert out["assignments"] == want and out["trees"] == os.path.realpath(path)
