"""-m gpu: the evolve / multievolve command lines end to end on the bundled example: output files and formats, status lines,
checkpoint + resume after SIGTERM (exit code 3, evolve.py:434-472), chain merging."""
import os
import re
import signal
import subprocess
import sys
import time
import zipfile

import pytest

from tests import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SSM = os.path.join(cases.GOLDEN, "ssm_data.txt")
CNV = os.path.join(cases.GOLDEN, "cnv_data.txt")
ENV = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))


def evolve(out, *extra, **kw):
    cmd = [sys.executable, "-m", "phylowgs_b200.host.evolve", "-O", out] + list(extra) + [SSM, CNV]
    return subprocess.run(cmd, env=ENV, cwd=ROOT, capture_output=True, text=True, timeout=600, **kw)


def tree_names(out):
    with zipfile.ZipFile(os.path.join(out, "trees.zip")) as z:
        assert z.testzip() is None
        return sorted(z.namelist())


def test_evolve_cli_outputs(gpu_lib, tmp_path):
    out = str(tmp_path / "run")
    r = evolve(out, "-B", "6", "-s", "9", "-r", "4", "-i", "500")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    status = [l for l in r.stdout.splitlines() if "iteration=" in l]
    assert len(status) == 15
    last = dict(f.split("=", 1) for f in status[-1].split("] ", 1)[1].split(" "))
    assert set(last) >= {"iteration", "trees_sampled", "total_trees", "llh", "nodes", "mh_acc", "dp_alpha", "dp_gamma", "alpha_decay"}
    assert last["trees_sampled"] == "14" and last["total_trees"] == "15"
    names = tree_names(out)
    assert sum(n.startswith("tree_") for n in names) == 9 and sum(n.startswith("burnin_") for n in names) == 6
    assert {"cnv_logical_physical_mapping.json", "params.json"} <= set(names)
    assert open(os.path.join(out, "random_seed.txt")).read().strip() == "4"
    rows = open(os.path.join(out, "mcmc_samples.txt")).read().strip().split("\n")
    assert rows[0] == "Iteration\tLLH\tTime" and len(rows) == 16
    for fn in ("state.initial.pickle", "state.last.pickle"):
        assert os.path.exists(os.path.join(out, fn))
    assert "Run succeeded." in r.stdout
    # the members are protocol-2 pickles a plain reader loads (util2.py:264-343): consistent trees, data attached
    import pickle
    from phylowgs_b200.host.evolve import install_pickle_aliases
    install_pickle_aliases()
    with zipfile.ZipFile(os.path.join(out, "trees.zip")) as z:
        for name in [n for n in names if n.startswith("tree_")][:3]:
            blob = z.read(name)
            assert blob[:2] == b"\x80\x02"
            t = pickle.loads(blob)
            nodes = t.get_nodes()
            assert sum(len(nd.data) for nd in nodes) == t.num_data == len(t.data) == 12
            for n, d in enumerate(t.data):
                assert d.tssb is t and d.node is t.assignments[n] and n in d.node.data
            assert [c.id for c, _, _ in t.data[2].cnv] == ["c0"] and t.data[2].cnv[0][0] is t.data[11]


def test_evolve_resumes_after_sigterm_and_reproduces_the_uninterrupted_chain(gpu_lib, tmp_path):
    args = ["-B", "10", "-s", "30", "-r", "8", "-i", "300", "-S", "5"]
    ref = str(tmp_path / "straight")
    assert evolve(ref, *args).returncode == 0
    out = str(tmp_path / "interrupted")
    cmd = [sys.executable, "-m", "phylowgs_b200.host.evolve", "-O", out] + args + [SSM, CNV]
    p = subprocess.Popen(cmd, env=ENV, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    seen = 0
    for line in p.stdout:                                   # let it write at least one checkpoint, then interrupt
        if "iteration=" in line:
            seen += 1
            if seen == 13:
                p.send_signal(signal.SIGTERM)
                break
    p.stdout.read()
    assert p.wait(timeout=120) == 3                         # interrupted runs exit with 3 (evolve.py:464)
    assert os.path.exists(os.path.join(out, "state.last.pickle"))
    r = evolve(out)                                          # CLI ignored on resume except -O (evolve.py:347-349)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Resuming existing run" in r.stdout
    assert tree_names(out) == tree_names(ref)                # member names carry index and LLH: the chain is the same chain


def test_multievolve_two_chains_on_one_gpu(gpu_lib, tmp_path):
    out = str(tmp_path / "chains")
    cmd = [sys.executable, "-m", "phylowgs_b200.host.multievolve", "-n", "2", "-r", "1", "2", "-O", out, "--ssms", SSM, "--cnvs", CNV,
           "-B", "4", "-s", "6", "-i", "300"]
    r = subprocess.run(cmd, env=ENV, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    for i in range(2):
        assert sum(n.startswith("tree_") for n in tree_names(os.path.join(out, "chain_%d" % i))) == 6
    merged = tree_names(out)
    n_trees = sum(n.startswith("tree_") for n in merged)
    assert n_trees in (6, 12) and "params.json" in merged and not any(n.startswith("burnin") for n in merged)
    assert re.search(r"Including chains", r.stdout)


def test_multievolve_in_process_chains_equal_the_per_process_chains(gpu_lib, tmp_path):
    """--in-process runs the chains as threads of one process (one CUDA context): same seeds -> the same chains, file for
    file, as one evolve process per chain."""
    base = [sys.executable, "-m", "phylowgs_b200.host.multievolve", "-n", "3", "-r", "5", "6", "7", "--ssms", SSM, "--cnvs", CNV,
            "-B", "5", "-s", "8", "-i", "300"]
    outs = {}
    for mode, extra in (("proc", []), ("thread", ["--in-process"])):
        outs[mode] = str(tmp_path / mode)
        r = subprocess.run(base + ["-O", outs[mode]] + extra, env=ENV, cwd=ROOT, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
        assert re.search(r"Including chains", r.stdout)
    assert "chain=2 status=done exit_code=0" in r.stdout
    for i in range(3):
        a, b = (os.path.join(outs[m], "chain_%d" % i) for m in ("proc", "thread"))
        assert tree_names(a) == tree_names(b)
        assert open(os.path.join(a, "random_seed.txt")).read() == open(os.path.join(b, "random_seed.txt")).read()
        rows = open(os.path.join(b, "mcmc_samples.txt")).read().strip().split("\n")
        assert len(rows) == 14
    assert tree_names(outs["proc"]) == tree_names(outs["thread"])
    assert not [f for f in os.listdir(ROOT) if f in ("trees.zip", "mcmc_samples.txt", "random_seed.txt")]   # nothing leaks into cwd
