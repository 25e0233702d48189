"""-m "not gpu": the evolve driver itself (CLI parsing, files, status lines, background checkpoint writes, tree pickles, resume)
with the CPU oracle plugged in where the product creates its GpuBackend -- the same test body as tests/test_gpu_cli.py runs with
the device.  TEST ONLY: the product never routes through the oracle (tests/test_abi.py)."""
import os
import pickle
import threading
import zipfile

import pytest

from phylowgs_b200.host import evolve as E
from tests import cases
from tests.oracle_backend import OracleBackend

SSM = os.path.join(cases.GOLDEN, "ssm_data.txt")
CNV = os.path.join(cases.GOLDEN, "cnv_data.txt")


@pytest.fixture()
def oracle_backend(monkeypatch):
    monkeypatch.setattr(E, "GpuBackend", lambda codes, device=0: OracleBackend(codes))


def run(out, *extra):
    safe, ok = threading.Event(), threading.Event()
    config = {"tmp_dir": None}
    E.run(safe, ok, config, ["-O", out] + list(extra) + [SSM, CNV])
    E.remove_tmp_files(config["tmp_dir"])
    return ok.is_set()


def names(out):
    with zipfile.ZipFile(os.path.join(out, "trees.zip")) as z:
        assert z.testzip() is None
        return sorted(z.namelist())


def test_driver_outputs(oracle_backend, tmp_path, capsys):
    out = str(tmp_path / "run")
    assert run(out, "-B", "4", "-s", "7", "-r", "4", "-i", "120", "-S", "3")
    status = [l for l in capsys.readouterr().out.splitlines() if "iteration=" in l]
    assert len(status) == 11 and "trees_sampled=10 total_trees=11" in status[-1]
    got = names(out)
    assert sum(n.startswith("tree_") for n in got) == 7 and sum(n.startswith("burnin_") for n in got) == 4
    assert {"cnv_logical_physical_mapping.json", "params.json"} <= set(got)
    rows = open(os.path.join(out, "mcmc_samples.txt")).read().strip().split("\n")
    assert rows[0] == "Iteration\tLLH\tTime" and [r.split("\t")[0] for r in rows[1:]] == [str(i) for i in range(-4, 7)]
    assert open(os.path.join(out, "random_seed.txt")).read().strip() == "4"
    E.install_pickle_aliases()
    with zipfile.ZipFile(os.path.join(out, "trees.zip")) as z:
        for n in got:
            if n.startswith("tree_"):
                t = pickle.loads(z.read(n))                      # plain reader (util2.py:264-343)
                assert float(n.split("_")[2]) < 0 and t.num_data == 12
                assert all(d.node is t.assignments[i] and d.tssb is t for i, d in enumerate(t.data))
    with open(os.path.join(out, "state.last.pickle"), "rb") as f:
        state = pickle.load(f)
    assert state["last_iteration"] == 6 and state["tssb"].num_data == 12
    assert not os.path.exists(os.path.join(out, "state.last.pickle.backup"))


def test_resume_continues_the_same_chain(oracle_backend, tmp_path):
    args = ["-B", "3", "-s", "9", "-r", "8", "-i", "100", "-S", "2"]
    straight = str(tmp_path / "straight")
    assert run(straight, *args)
    cut = str(tmp_path / "cut")
    assert run(cut, "-B", "3", "-s", "4", "-r", "8", "-i", "100", "-S", "2")      # a shorter run ...
    with open(os.path.join(cut, "state.last.pickle"), "rb") as f:
        state = pickle.load(f)
    state["num_samples"] = 9                                                        # ... whose state is asked for 9 samples after all
    import numpy as np
    trace = np.zeros((9, 1))
    trace[:4] = state["cd_llh_traces"]
    state["cd_llh_traces"] = trace
    with open(os.path.join(cut, "state.last.pickle"), "wb") as f:
        pickle.dump(state, f, protocol=2)
    assert run(cut)                                                                 # resumes: CLI ignored but -O (evolve.py:347-349)
    assert names(cut) == names(straight)                                            # member names carry index and LLH
    rows = lambda d: [r.split("\t")[:2] for r in open(os.path.join(d, "mcmc_samples.txt")).read().strip().split("\n")]   # noqa: E731
    assert rows(cut) == rows(straight)


def test_sync_writes_and_plain_pickle_switches_give_the_same_files(oracle_backend, tmp_path, monkeypatch):
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    args = ["-B", "2", "-s", "4", "-r", "5", "-i", "80", "-S", "2"]
    assert run(a, *args)
    monkeypatch.setenv("PWGS_SYNC_WRITES", "1")
    monkeypatch.setenv("PWGS_PLAIN_PICKLE", "1")
    assert run(b, *args)
    assert names(a) == names(b)


def test_multievolve_in_process_with_the_oracle(oracle_backend, tmp_path, capsys):
    """Chains as threads of one process (multievolve --in-process): per-chain directories, status relay, merged trees.zip --
    and the same chains as the single-chain driver gives for the same seeds."""
    from phylowgs_b200.host import multievolve as M
    out = str(tmp_path / "chains")
    M.main(["-n", "2", "-r", "5", "6", "--gpus", "0", "--in-process", "-O", out, "--ssms", SSM, "--cnvs", CNV,
            "-B", "2", "-s", "4", "-i", "80", "-S", "2"])
    text = capsys.readouterr().out
    assert "chain=1 status=done exit_code=0" in text and "Including chains" in text
    single = str(tmp_path / "single")
    assert run(single, "-B", "2", "-s", "4", "-r", "6", "-i", "80", "-S", "2")
    assert names(os.path.join(out, "chain_1")) == names(single)
    merged = names(out)
    assert sum(n.startswith("tree_") for n in merged) in (4, 8) and not any(n.startswith("burnin") for n in merged)
