"""host/write_results.py (write_results.py + pwgsresults/) against the reference's own classes.

  * frozen: tests/golden/results_*.zip (hand-shaped trees: superclones, multiprimary roots, tiny populations, chains, random
    branchings; written by the product's pickler) with the reference's outputs for several option sets in
    tests/golden/results_vectors.json (tests/golden/make_results_golden.py ran the reference's ResultGenerator / ResultMunger /
    IndexCalculator / calc_tree_densities under Python 3 through oracle/ref_py3.py);
  * live, where /root/reference is present: the same comparison on freshly drawn forests and on the archive of a real chain.
Integers, ids and structures must be equal; floating-point values to 1e-12 (they are the same operations in the same order)."""
import gzip
import json
import os
import threading
import zipfile

import numpy as np
import pytest

from oracle import ref_py3
from phylowgs_b200.host import evolve as E
from phylowgs_b200.host import write_results as W
from tests import cases, results_util as U
from tests.oracle_backend import OracleBackend

GOLDEN = cases.GOLDEN
OPTION_SETS = {
    "defaults": dict(),
    "names_min3": dict(include_ssm_names=True, min_ssms=3),
    "keep_all": dict(keep_superclones=True, include_multiprimary=True, min_ssms=0.0),
    "frac20": dict(min_ssms=0.2, include_multiprimary=True),
    "count1_strict": dict(min_ssms=1, max_multiprimary=0.9),
}


def ours(archive, **kw):
    """Our three outputs in the shape oracle/ref_py3.reference_results gives the reference's."""
    trees, mutlist, params = W.summarize(archive, **kw)
    body, indices = {}, {}
    for idx, s in trees.items():
        indices[idx] = s.indices()
        body[idx] = s.as_json()
        body[idx]["linearity_index"], body[idx]["branching_index"], body[idx]["clustering_index"] = indices[idx]
    out = {"summaries": {"params": params, "trees": body, "tree_densities": W.tree_densities(indices)}, "mutlist": mutlist,
           "mutass": {idx: {"mut_assignments": s.mutass} for idx, s in trees.items()}}
    return json.loads(json.dumps(out, default=W._plain))


def same(a, b, path="$"):
    if isinstance(a, dict):
        assert isinstance(b, dict) and sorted(a) == sorted(b), (path, sorted(a), sorted(b) if isinstance(b, dict) else b)
        for k in a:
            same(a[k], b[k], path + "." + str(k))
    elif isinstance(a, list):
        assert isinstance(b, list) and len(a) == len(b), (path, a, b)
        for i, (x, y) in enumerate(zip(a, b)):
            same(x, y, "%s[%d]" % (path, i))
    elif isinstance(a, float) or isinstance(b, float):
        assert abs(a - b) <= 1e-12 * max(1.0, abs(a), abs(b)), (path, a, b)
    else:
        assert a == b and type(a) is type(b), (path, a, b)


VECTORS = os.path.join(GOLDEN, "results_vectors.json")


@pytest.mark.parametrize("archive", ["results_bundled.zip", "results_sim150.zip"])
def test_outputs_equal_the_frozen_reference_outputs(archive):
    with open(VECTORS) as f:
        frozen = json.load(f)[archive]
    for name, kw in OPTION_SETS.items():
        if "error" in frozen[name]:
            with pytest.raises(Exception, match="multiprimary"):
                ours(os.path.join(GOLDEN, archive), **kw)
        else:
            same(frozen[name], ours(os.path.join(GOLDEN, archive), **kw), "$." + name)


def test_too_many_multiprimary_trees_is_an_error(tmp_path):
    ssm, cnv = os.path.join(GOLDEN, "ssm_data.txt"), os.path.join(GOLDEN, "cnv_data.txt")
    trees = []
    for i in range(4):
        t, mapping = U.build_tree(ssm, cnv, [-1, 0, 0], [1.0, 0.5, 0.3] if i else [1.0, 0.6, 0.2], 1 + np.arange(12) % 2, seed=i)
        trees.append((t, -100.0 - i))
    t, mapping = U.build_tree(ssm, cnv, [-1, 0, 1], [1.0, 0.5, 0.3], 1 + np.arange(12) % 2, seed=9)
    trees.append((t, -90.0))
    path = str(tmp_path / "trees.zip")
    U.write_archive(path, trees, mapping)
    with pytest.raises(Exception, match="80% of trees are multiprimary"):
        W.summarize(path)
    kept, _, params = W.summarize(path, max_multiprimary=0.9)
    assert sorted(kept) == [4] and params == {}                             # survivors keep their index; no params.json in the archive
    assert sorted(W.summarize(path, include_multiprimary=True)[0]) == [0, 1, 2, 3, 4]


def test_command_line_writes_the_three_files(tmp_path, capsys):
    out = [str(tmp_path / n) for n in ("x.summ.json.gz", "x.muts.json.gz", "x.mutass.zip")]
    assert W.main(["--include-ssm-names", "tumour-7", os.path.join(GOLDEN, "results_bundled.zip")] + out) == 0
    summ = json.loads(gzip.open(out[0]).read())
    muts = json.loads(gzip.open(out[1]).read())
    assert summ["dataset_name"] == muts["dataset_name"] == "tumour-7"
    assert set(summ) == {"dataset_name", "params", "trees", "tree_densities"} and set(summ["trees"]) == set(summ["tree_densities"])
    assert set(muts) == {"ssms", "cnvs", "dataset_name"} and muts["ssms"]["s0"]["name"] and muts["cnvs"]["c0"]["physical_cnvs"]
    with zipfile.ZipFile(out[2]) as z:
        assert sorted(z.namelist()) == sorted("%s.json" % i for i in summ["trees"])
        for name in z.namelist():
            m = json.loads(z.read(name))
            tree = summ["trees"][name[:-5]]
            assert m["dataset_name"] == "tumour-7"
            for pop, held in m["mut_assignments"].items():
                assert len(held["ssms"]) == tree["populations"][pop]["num_ssms"] and len(held["cnvs"]) == tree["populations"][pop]["num_cnvs"]
            assert sum(len(h["ssms"]) for h in m["mut_assignments"].values()) == 11
    assert "Superclone" in capsys.readouterr().out                          # the reference reports its merges on stdout
    with pytest.raises(SystemExit):
        W.main(["--max-multiprimary", "1.5", "x", "a", "b", "c", "d"])


def test_indices_against_the_pairwise_definition():
    """index_calculator.py:11-49 spelled out over all pairs of populations, on a tree with cousins at several depths."""
    s = W.TreeSummary(-1.0, {i: {"cellular_prevalence": [1.0 / (i + 1)], "num_ssms": n, "num_cnvs": 0} for i, n in enumerate([0, 5, 3, 7, 2, 4])},
                      {0: [1], 1: [2, 4], 2: [3], 4: [5]}, {})
    anc = {(0, k) for k in range(1, 6)} | {(1, 2), (1, 3), (1, 4), (1, 5), (2, 3), (4, 5)}
    n = [0, 5, 3, 7, 2, 4]
    tot = sum(n)
    lin = 2 * sum(n[a] * n[b] for a, b in anc) / float(tot * (tot - 1))
    cousins = [(a, b) for a in range(6) for b in range(6) if a != b and (a, b) not in anc and (b, a) not in anc]
    bra = sum(n[a] * n[b] for a, b in cousins) / float(tot * (tot - 1))
    clu = sum(v * (v - 1) for v in n) / float(tot * (tot - 1))
    assert np.allclose(s.indices(), (lin, bra, clu), rtol=0, atol=1e-15)


@pytest.fixture()
def chain_archive(tmp_path, monkeypatch):
    monkeypatch.setattr(E, "GpuBackend", lambda codes, device=0: OracleBackend(codes))
    out = str(tmp_path / "run")
    safe, ok = threading.Event(), threading.Event()
    config = {"tmp_dir": None}
    E.run(safe, ok, config, ["-O", out, "-B", "4", "-s", "12", "-r", "5", "-i", "200", "-S", "3",
                             os.path.join(GOLDEN, "ssm_data.txt"), os.path.join(GOLDEN, "cnv_data.txt")])
    E.remove_tmp_files(config["tmp_dir"])
    assert ok.is_set()
    return os.path.join(out, "trees.zip")


@pytest.mark.skipif(not ref_py3.available(), reason="needs the reference sources (build container only)")
def test_live_against_the_reference_on_a_chain_and_on_fresh_forests(chain_archive, tmp_path):
    try:
        for kw in (dict(include_multiprimary=True), dict(min_ssms=2, include_multiprimary=True, keep_superclones=True)):
            same(ref_py3.reference_results(chain_archive, **kw), ours(chain_archive, **kw))
        ssm = os.path.join(GOLDEN, "emptysim.3.20.10.ssm.txt")
        empty = str(tmp_path / "no_cnvs.txt")
        open(empty, "w").close()
        for seed in (11, 12, 13):
            trees, mapping = U.shaped_forest(ssm, empty, 20, seed)
            path = str(tmp_path / ("forest%d.zip" % seed))
            U.write_archive(path, trees, mapping, params={"seed": seed})
            for kw in (dict(), dict(min_ssms=0.15), dict(min_ssms=4, keep_superclones=True)):
                same(ref_py3.reference_results(path, **kw), ours(path, **kw), "$.seed%d" % seed)
    finally:
        E.install_pickle_aliases()
