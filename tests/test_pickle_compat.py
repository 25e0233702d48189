"""-m "not gpu": what the evolve driver writes must be readable by the reference's readers (SURVEY.md 8b "Output compat":
util2.TreeReader util2.py:264-343 under Python 2 / numpy 1, pwgsresults/result_generator.py:11-35).

Three levels: (1) the pickle stream itself — protocol <= 2 opcodes only and a whitelist of GLOBALs that exist under Python 2 with
numpy 1.x (numpy >= 2 would name `numpy._core.multiarray`, which numpy 1 does not have); (2) the archive layout (member names,
extra files); (3) where /root/reference is present, the reference's own TreeReader and ResultGenerator — its code, executed by
oracle/ref_py3.py, with ITS classes on the receiving side — read the archive and summarise it."""
import json
import os
import pickle
import pickletools
import sys
import threading
import zipfile

import numpy as np
import pytest

from oracle import ref_py3
from phylowgs_b200.host import evolve as E
from phylowgs_b200.host import fastpickle
from tests import cases
from tests.oracle_backend import OracleBackend
from tests.test_fastpickle import make_tree

SSM = os.path.join(cases.GOLDEN, "ssm_data.txt")
CNV = os.path.join(cases.GOLDEN, "cnv_data.txt")

# module, name pairs a Python-2 / numpy-1 unpickler can resolve
ALLOWED_GLOBALS = {"tssb TSSB", "alleles alleles", "node Node", "data Datum", "numpy ndarray", "numpy dtype",
                   "numpy.core.multiarray _reconstruct", "numpy.core.multiarray scalar", "__builtin__ set", "__builtin__ bytes",
                   "_codecs encode"}


def stream_facts(blob):
    globs, max_proto, names = set(), 0, set()
    for op, arg, _ in pickletools.genops(blob):
        max_proto = max(max_proto, op.proto)
        names.add(op.name)
        if op.name == "GLOBAL":
            globs.add(arg)
    return globs, max_proto, names


@pytest.mark.parametrize("plain", [False, True], ids=["fastpickle", "compat_dumps"])
def test_tree_pickle_names_only_what_python2_numpy1_can_import(plain):
    t = make_tree()
    blob = fastpickle.compat_dumps(t) if plain else fastpickle.dumps(t)
    globs, max_proto, names = stream_facts(blob)
    assert globs <= ALLOWED_GLOBALS, globs - ALLOWED_GLOBALS
    assert "numpy.core.multiarray _reconstruct" in globs          # arrays are in there, under their numpy-1 name
    assert max_proto <= 2 and "STACK_GLOBAL" not in names and "FRAME" not in names
    back = pickle.loads(blob)                                      # and this numpy reads it too
    assert np.array_equal(back.get_nodes()[1].pi, t.get_nodes()[1].pi)
    assert back.get_nodes()[1].pi.flags.writeable


def test_state_pickle_names_only_what_python2_numpy1_can_import():
    t = make_tree()
    state = dict(tssb=t, rand_state=np.random.RandomState(3).get_state(), cd_llh_traces=np.zeros((4, 1)), mh_std=np.float64(100.0),
                 last_iteration=2, burnin=1, num_samples=4)
    for blob in (E.StateManager.dumps(state), fastpickle.compat_dumps(state)):
        globs, max_proto, _ = stream_facts(blob)
        assert globs <= ALLOWED_GLOBALS, globs - ALLOWED_GLOBALS
        assert {"numpy.core.multiarray _reconstruct", "numpy.core.multiarray scalar"} <= globs
        assert max_proto <= 2
        back = pickle.loads(blob)
        assert back["mh_std"] == 100.0 and np.array_equal(back["rand_state"][1], state["rand_state"][1])


@pytest.fixture()
def archive(tmp_path, monkeypatch):
    monkeypatch.setattr(E, "GpuBackend", lambda codes, device=0: OracleBackend(codes))
    out = str(tmp_path / "run")
    safe, ok = threading.Event(), threading.Event()
    config = {"tmp_dir": None}
    E.run(safe, ok, config, ["-O", out, "-B", "3", "-s", "6", "-r", "11", "-i", "150", "-S", "2", SSM, CNV])
    E.remove_tmp_files(config["tmp_dir"])
    assert ok.is_set()
    return os.path.join(out, "trees.zip")


def test_every_archive_member_is_python2_numpy1_material(archive):
    with zipfile.ZipFile(archive) as z:
        members = [n for n in z.namelist() if n.startswith(("tree_", "burnin_"))]
        assert len(members) == 9
        for n in members:
            globs, max_proto, _ = stream_facts(z.read(n))
            assert globs <= ALLOWED_GLOBALS, (n, globs - ALLOWED_GLOBALS)
            assert max_proto <= 2


@pytest.mark.skipif(not ref_py3.available(), reason="needs the reference sources (build container only)")
def test_reference_readers_consume_the_archive(archive):
    """util2.TreeReader + pwgsresults.ResultGenerator — the reference's code and the reference's classes — on our trees.zip."""
    aliases = {n: sys.modules.get(n) for n in ("tssb", "alleles", "node", "data")}
    R = ref_py3.load()                                  # installs the reference's tssb / alleles / node / data in sys.modules
    try:
        reader = R.util2.TreeReader(archive)
        assert reader.num_trees() == 6
        seen = 0
        for idx, llh, tree in reader.load_trees_and_metadata(remove_empty_vertices=True):
            assert type(tree).__module__ == "tssb" and type(tree) is R.tssb.TSSB
            assert type(tree.root["node"]) is R.alleles.alleles and type(tree.data[0]) is R.data.Datum
            weights, nodes = tree.get_mixture()                                    # tssb.py:389-402 on the loaded object
            assert abs(sum(weights) - 1.0) < 1e-6 or sum(weights) < 1.0
            assert sum(len(nd.get_data()) for nd in nodes) == 12
            # the reference's own likelihood code runs on the loaded tree and agrees with the LLH in the member's name
            R.util2.set_node_height(tree), R.util2.set_path_from_root_to_node(tree), R.util2.map_datum_to_node(tree)
            seen += 1
        assert seen == 6
        burnin = [i for i, _ in reader.load_trees_and_burnin()]
        assert burnin == [-3, -2, -1, 0, 1, 2, 3, 4, 5]
        reader.close()
        sys.path.insert(0, ref_py3.REF)
        try:
            from pwgsresults.result_generator import ResultGenerator
            summaries, mutlist, mutass, params = ResultGenerator().generate(archive, include_ssm_names=True)
        finally:
            sys.path.remove(ref_py3.REF)
            for m in [m for m in sys.modules if m.startswith("pwgsresults")]:
                del sys.modules[m]
        assert sorted(summaries) == list(range(6))
        assert len(mutlist["ssms"]) == 11 and len(mutlist["cnvs"]) == 1
        assert mutlist["cnvs"]["c0"]["ssms"] and mutlist["ssms"]["s0"]["name"]
        for idx, s in summaries.items():
            n_ssms = sum(p["num_ssms"] for p in s["populations"].values())
            assert n_ssms == 11 and len(s["populations"][0]["cellular_prevalence"]) == 5
            assert set(mutass[idx]) <= set(s["populations"])
        json.dumps(summaries)                                                       # plain Python values all the way down
    finally:
        R.release()
        for n, m in aliases.items():
            if m is not None:
                sys.modules[n] = m
            else:
                sys.modules.pop(n, None)
        E.install_pickle_aliases()


@pytest.mark.skipif(not ref_py3.available(), reason="needs the reference sources (build container only)")
def test_reference_complete_data_log_likelihood_of_a_loaded_tree_matches_the_member_name(archive):
    """The LLH in a member's name (computed on the device / oracle) is what the reference's own tssb.py + data.py compute on the
    tree they unpickle from that member: pickled sticks, pi, params and assignments carry everything."""
    aliases = {n: sys.modules.get(n) for n in ("tssb", "alleles", "node", "data")}
    R = ref_py3.load()
    try:
        reader = R.util2.TreeReader(archive)
        for idx, llh, tree in reader.load_trees_and_metadata():
            for dat in tree.data:
                assert dat.tssb is tree
            got = tree.complete_data_log_likelihood()
            assert abs(got - llh) <= 1e-9 * abs(llh)
        reader.close()
    finally:
        R.release()
        for n, m in aliases.items():
            if m is not None:
                sys.modules[n] = m
            else:
                sys.modules.pop(n, None)
        E.install_pickle_aliases()
