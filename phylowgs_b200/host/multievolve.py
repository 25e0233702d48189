"""Multi-chain driver with the reference's CLI and outputs (multievolve.py:20-323), plus chain -> GPU placement.

    python -m phylowgs_b200.host.multievolve -n 8 [-r s0 s1 ...] [-I 1.1] [-O chains] --ssms ssm.txt --cnvs cnv.txt [evolve options]

Chains share nothing (multievolve.py:85-90), so chain i runs as its own `evolve` process pinned to GPU i mod n_gpus through
PWGS_DEVICE — no collective, no NCCL.  With --in-process the chains are threads of this process instead (one device context
and stream per chain; ctypes releases the GIL during every library call, so one chain's host work overlaps the others' kernels
and native sweeps): several chains per GPU then share one CUDA context instead of time-slicing between processes.  Afterwards the chains whose log-sum of tree likelihoods is within the inclusion factor
of the best one are merged into <output-dir>/trees.zip with re-indexed tree members (multievolve.py:215-313).
"""
import argparse
import hashlib
import os
import queue
import re
import subprocess
import sys
import threading
import time
import zipfile
from collections import defaultdict

import numpy as np

from .evolve import logmsg


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Concurrently run multiple MCMC chains of PhyloWGS. All options that evolve accepts "
                                            "may also be specified here.", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("-n", "--num-chains", dest="num_chains", default=4, type=int, help="Number of chains to run concurrently")
    p.add_argument("-r", "--random-seeds", dest="random_seeds", type=int, nargs="+",
                   help="Space-separated random seeds, one for each chain")
    p.add_argument("-I", "--chain-inclusion-factor", dest="chain_inclusion_factor", default=1.1, type=float,
                   help="A chain is merged if the log-sum of its tree likelihoods is >= this factor times the best chain's "
                        "(likelihoods are negative); inf includes all chains, 1 only the best")
    p.add_argument("-O", "--output-dir", dest="output_dir", default="chains", help="Directory for the per-chain results")
    p.add_argument("--ssms", dest="ssm_file", required=True, help="File listing SSMs")
    p.add_argument("--cnvs", dest="cnv_file", required=True, help="File listing CNVs")
    p.add_argument("--gpus", dest="gpus", default=None,
                   help="Comma-separated device indices to spread the chains over (default: all visible devices)")
    p.add_argument("--in-process", dest="in_process", action="store_true",
                   help="Run the chains as threads of this process (one CUDA context per GPU) instead of one process per chain")
    known, other = p.parse_known_args(argv)
    return vars(known), other


def visible_devices(spec=None):
    if spec:
        return [int(x) for x in spec.split(",")]
    from .. import device_count
    return list(range(max(device_count(), 1)))


def chain_device(chain_index, devices):
    """Chain i -> devices[i mod n]: round-robin, so 8 chains on 8 GPUs get one GPU each and 16 chains two per GPU."""
    return devices[chain_index % len(devices)]


def run_chain(chain_index, ssm_fn, cnv_fn, working_dir, output_dir, seeds, other_args, device):
    cmd = [sys.executable, "-m", "phylowgs_b200.host.evolve", "--output-dir", output_dir]
    if seeds is not None:
        cmd += ["--random-seed", str(seeds[chain_index])]
    cmd += [ssm_fn, cnv_fn] + list(other_args)
    env = dict(os.environ, PWGS_DEVICE=str(device))
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    logmsg("Starting chain %s on GPU %s" % (chain_index, device))
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, bufsize=1, universal_newlines=True,
                            close_fds=True, cwd=working_dir, env=env)


def parse_status(line):
    return dict(f.split("=", 1) for f in line.split(" "))


def watch_chains(processes, poll=1.0):
    """Relay the chains' output and a one-line progress report per chain until all have exited (multievolve.py:151-213)."""
    n = len(processes)
    queues = []
    for proc in processes:
        q = queue.Queue()

        def pump(out=proc.stdout, q=q):
            for line in iter(out.readline, ""):
                q.put(line)
            out.close()
        threading.Thread(target=pump, daemon=True).start()
        queues.append(q)
    status = {i: {"status": "initializing"} for i in range(n)}
    while any(s["status"] != "done" for s in status.values()):
        for i, (proc, q) in enumerate(zip(processes, queues)):
            if status[i]["status"] == "done":
                continue
            code = proc.poll()
            lines = []
            while True:
                try:
                    lines.append(q.get(timeout=0.05).strip())
                except queue.Empty:
                    break
            for line in lines:
                if re.match(r"^\[\d{4}-", line):
                    line = line[line.index("]") + 1:].strip()
                if line.startswith("iteration="):
                    st = parse_status(line)
                    st["status"] = "running"
                    st["percent_complete"] = "{:.2f}%".format(100 * float(st["trees_sampled"]) / float(st["total_trees"]))
                    status[i] = st
                elif line:
                    logmsg("chain={} {}".format(i, line))
            if code is not None:
                status[i] = {"status": "done", "exit_code": code}
        for i in sorted(status):
            keys = {"running": ("trees_sampled", "total_trees", "percent_complete"), "done": ("exit_code",)}.get(status[i]["status"], ())
            logmsg("chain={} {}".format(i, " ".join("{}={}".format(k, status[i][k]) for k in ("status",) + keys)))
        time.sleep(poll)
    return [status[i].get("exit_code") for i in range(n)]


def run_chains_in_process(chain_dirs, ssm_fn, cnv_fn, seeds, other_args, devices, poll=1.0):
    """All chains as threads of this process, each driving evolve.run with its own directory, device and status callback.
    -> exit codes like watch_chains (0 succeeded, 1 failed).  SIGTERM / SIGINT wait until every chain is between file writes,
    then exit 3 (each chain resumes from its state.last.pickle on the next run)."""
    import signal
    from . import evolve
    n = len(chain_dirs)
    status = {i: {"status": "initializing"} for i in range(n)}
    chains = []
    # a chain coming back from a library call must get the GIL back from whichever chain is running Python; the default
    # 5 ms hand-over interval would add up to more than the calls themselves
    sys.setswitchinterval(float(os.environ.get("PWGS_SWITCH_INTERVAL", "2e-4")))
    for i, d in enumerate(chain_dirs):
        argv = ["--output-dir", d]
        if seeds is not None:
            argv += ["--random-seed", str(seeds[i])]
        argv += [ssm_fn, cnv_fn] + list(other_args)

        def on_status(st, i=i):
            st = dict(st, status="running")
            st["percent_complete"] = "{:.2f}%".format(100 * float(st["trees_sampled"]) / float(st["total_trees"]))
            status[i] = st
        safe, ok = threading.Event(), threading.Event()
        safe.set()
        config = {"tmp_dir": None, "device": chain_device(i, devices), "log_prefix": "chain=%d " % i, "on_status": on_status}

        def body(safe=safe, ok=ok, config=config, argv=argv):
            try:
                evolve.run(safe, ok, config, argv)
            except SystemExit:
                pass
            except Exception:
                import traceback
                logmsg(config["log_prefix"] + "failed:", sys.stderr)
                traceback.print_exc()
        logmsg("Starting chain %s on GPU %s (in-process)" % (i, config["device"]))
        th = threading.Thread(target=body, daemon=True)
        chains.append((th, safe, ok, config))

    def on_signal(signo, _frame):
        logmsg("Signal %s received." % signo, sys.stderr)
        for _, safe, _, config in chains:
            safe.wait()
            if config.get("writer") is not None:
                config["writer"].stop_and_drain()
        for _, _, _, config in chains:
            evolve.remove_tmp_files(config["tmp_dir"])
        logmsg("Exiting now.")
        os._exit(3)             # the chain threads are mid-iteration by design; nothing is being written (all safe events held)
    previous = {}
    if threading.current_thread() is threading.main_thread():
        for sig in (signal.SIGTERM, signal.SIGINT):
            previous[sig] = signal.signal(sig, on_signal)
    for th, _, _, _ in chains:
        th.start()
    while any(th.is_alive() for th, _, _, _ in chains):
        time.sleep(poll)
        for i, (th, _, ok, _) in enumerate(chains):
            if not th.is_alive():
                status[i] = {"status": "done", "exit_code": 0 if ok.is_set() else 1}
            st = status[i]
            keys = {"running": ("trees_sampled", "total_trees", "percent_complete"), "done": ("exit_code",)}.get(st["status"], ())
            logmsg("chain={} {}".format(i, " ".join("{}={}".format(k, st[k]) for k in ("status",) + keys)))
    for _, _, _, config in chains:
        evolve.remove_tmp_files(config["tmp_dir"])
    for sig, handler in previous.items():
        signal.signal(sig, handler)
    return [0 if ok.is_set() else 1 for _, _, ok, _ in chains]


def logsumexp(x):
    x = np.asarray(x, dtype=float)
    m = np.max(x)
    return float(m + np.log(np.sum(np.exp(x - m))))


def determine_chains_to_merge(chain_dirs, factor):
    sums = []
    for d in chain_dirs:
        with zipfile.ZipFile(os.path.join(d, "trees.zip")) as z:
            sums.append(logsumexp([float(n.split("_")[-1]) for n in z.namelist() if n.startswith("tree")]))
    sums = np.array(sums)
    assert np.all(sums < 0)
    best = np.max(sums)
    inc = [(i, s) for i, s in enumerate(sums) if s >= factor * best]
    exc = [(i, s) for i, s in enumerate(sums) if s < factor * best]
    assert inc
    return inc, exc


def merge_best_chains(out_dir, chain_dirs, included, excluded):
    combined = os.path.join(out_dir, "trees.zip")
    if os.path.isfile(combined):
        logmsg("Merged trees.zip file already exists. To create a new merged trees.zip, remove the existing one first.")
        return
    logmsg("Including chains {}".format(" ".join("{}={}".format(i, s) for i, s in included)))
    logmsg("Excluding chains {}".format(" ".join("{}={}".format(i, s) for i, s in excluded)) if excluded else "Not excluding any chains")
    others, payload, index = defaultdict(dict), {}, 0
    with zipfile.ZipFile(combined, "w", compression=zipfile.ZIP_DEFLATED, allowZip64=True) as out:
        for i, _ in included:
            with zipfile.ZipFile(os.path.join(chain_dirs[i], "trees.zip")) as z:
                for fn in z.namelist():
                    if fn.startswith("tree"):
                        parts = fn.split("_")
                        parts[1] = str(index)
                        out.writestr("_".join(parts), z.read(fn))
                        index += 1
                    elif not fn.startswith("burnin"):
                        blob = z.read(fn)
                        others[fn][i] = hashlib.sha256(blob).hexdigest()
                        payload[fn] = blob
        for fn, digests in others.items():
            assert len(set(digests.values())) == 1, "non-tree member %s differs between chains" % fn
            out.writestr(fn, payload[fn])
    logmsg("Chain merging complete.")


def main(argv=None):
    args, evolve_args = parse_args(argv)
    args["output_dir"] = os.path.abspath(args["output_dir"])
    os.makedirs(args["output_dir"], exist_ok=True)
    if args["random_seeds"] is not None and len(args["random_seeds"]) != args["num_chains"]:
        raise ValueError("Must specify exactly one seed for each of %s chain(s). You specified %s seed(s)."
                         % (args["num_chains"], len(args["random_seeds"])))
    devices = visible_devices(args["gpus"])
    ssm, cnv = os.path.abspath(args["ssm_file"]), os.path.abspath(args["cnv_file"])
    dirs = [os.path.join(args["output_dir"], "chain_%d" % i) for i in range(args["num_chains"])]
    for d in dirs:
        os.makedirs(d, exist_ok=True)
    if args["in_process"]:
        codes = run_chains_in_process(dirs, ssm, cnv, args["random_seeds"], evolve_args, devices)
    else:
        procs = [run_chain(i, ssm, cnv, os.getcwd(), d, args["random_seeds"], evolve_args, chain_device(i, devices))
                 for i, d in enumerate(dirs)]
        codes = watch_chains(procs)
    if any(codes):
        logmsg("Chains %s failed; not merging." % [i for i, c in enumerate(codes) if c], sys.stderr)
        sys.exit(1)
    inc, exc = determine_chains_to_merge(dirs, args["chain_inclusion_factor"])
    merge_best_chains(args["output_dir"], dirs, inc, exc)


if __name__ == "__main__":
    main()
