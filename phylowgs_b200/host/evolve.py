"""Single-chain driver with the reference's CLI, files and status lines (evolve.py:31-497), Python 3, device-backed.

    python -m phylowgs_b200.host.evolve [-O DIR] [-B 1000] [-s 2500] [-i 5000] [-r SEED] [-b 100] [-S 10] [-t TMP] [-p params.json] ssm cnv

Writes trees.zip (members burnin_<i>_<llh> / tree_<i>_<llh> = protocol-2 pickles of the TSSB, util2.py:250-262),
mcmc_samples.txt, random_seed.txt, state.initial.pickle / state.last.pickle (+ .backup copies) exactly where the reference
does; resumes from state.last.pickle when it exists (CLI then ignored except -O, evolve.py:347-349); exit codes 0 / 1 / 3.
The device is chosen with PWGS_DEVICE (default 0) so the CLI surface is unchanged.
"""
import argparse
import csv
import json
import os
import pickle
import queue
import shutil
import signal
import sys
import tempfile
import threading
import time
import traceback
import zipfile
from collections import OrderedDict
from datetime import datetime

import numpy as np

from . import fastpickle, model
from .backend import GpuBackend
from .model import TSSB, Datum, alleles, boundbeta
from .params import get_c_fnames, metropolis

csv.field_size_limit(2147483647)        # .cnv rows can list thousands of SSMs (util2.py:13-18)


def logmsg(msg, fd=None):
    fd = fd or sys.stdout
    fd.write("[%s] %s\n" % (datetime.now(), msg))
    fd.flush()


# ------------------------------------------------------------------ inputs (util2.py:40-94)
def parse_physical_cnvs(text):
    out = []
    for item in text.split(";"):
        cnv = dict(f.split("=", 1) for f in item.split(","))
        for key in ("start", "end", "major_cn", "minor_cn"):
            cnv[key] = int(cnv[key])
        cnv["cell_prev"] = [float(c) for c in cnv["cell_prev"].split("|")]
        out.append(cnv)
    return out


def load_data(ssm_fn, cnv_fn):
    """-> (data in file order: SSMs then CNVs, n_ssms, n_cnvs, cnv_logical_physical_mapping).  The reference returns the data in
    Python-2 dict order of the ids (util2.py:94); file order is what mh.cpp indexes by (params.py:71-76) and what we use."""
    by_id = OrderedDict()
    with open(ssm_fn, newline="") as f:
        for row in csv.DictReader(f, delimiter="\t"):
            mu_r = float(row["mu_r"]) if "mu_r" in row else 0.0
            mu_v = float(row["mu_v"]) if "mu_v" in row else 0.0
            by_id[row["id"]] = Datum(row["gene"], row["id"], [int(x) for x in row["a"].split(",")],
                                     [int(x) for x in row["d"].split(",")], mu_r, mu_v)
    n_ssms = len(by_id)
    mapping = {}
    with open(cnv_fn, newline="") as f:
        for row in csv.DictReader(f, delimiter="\t"):
            cid = row["cnv"]
            mapping[cid] = parse_physical_cnvs(row["physical_cnvs"])
            cnv = Datum(cid, cid, [int(x) for x in row["a"].split(",")], [int(x) for x in row["d"].split(",")], 0.999, 0.5)
            by_id[cid] = cnv
            for tok in filter(None, (row.get("ssms") or "").split(";")):
                sid, c1, c2 = tok.split(",")
                by_id[sid].cnv.append((cnv, int(c1), int(c2)))
    data = list(by_id.values())
    for i, d in enumerate(data):
        d.index = i
    return data, n_ssms, len(data) - n_ssms, mapping


# ------------------------------------------------------------------ tree metadata (util2.py:99-115)
def set_node_height(tssb):
    stack = [(tssb.root["node"], 0)]
    while stack:
        nd, ht = stack.pop()
        nd.ht = ht
        stack.extend((c, ht + 1) for c in nd.children())


def set_path_from_root_to_node(tssb):
    for nd in tssb.get_nodes():
        nd.path = nd.get_ancestors()


def map_datum_to_node(tssb):
    for nd in tssb.get_nodes():
        for n in nd.data:
            tssb.data[n].node = nd


# ------------------------------------------------------------------ persistence (util2.py:157-262)
def install_pickle_aliases():
    """Pickle the model classes under the module names the reference's readers expect (tssb.TSSB, alleles.alleles, node.Node,
    data.Datum) and let this process unpickle them again."""
    import types
    for cls in (TSSB, alleles, model.Node, Datum):
        name = model.PICKLE_MODULES[cls.__name__]
        cls.__module__ = name
        mod = sys.modules.get(name)
        if mod is None:
            mod = types.ModuleType(name)
            sys.modules[name] = mod
        setattr(mod, cls.__name__, cls)


def tune_allocator():
    """Keep large blocks on the C heap instead of handing them back to the kernel after every use.  A tree sample at 50k SSMs
    allocates and frees several 8 MB blocks (the pickle, numpy temporaries); glibc serves those with mmap / munmap, so every one
    of them is page-faulted in again, 2-3 us per page in the virtualised hosts this runs on (8 MB: 5 ms).  With the thresholds
    raised the blocks are recycled (measured: pickle of a tree 19.5 -> 14 ms before the native join, 8 MB numpy fill 1.6 -> 0.6 ms).
    PWGS_NO_MALLOPT=1 leaves the allocator alone."""
    if os.environ.get("PWGS_NO_MALLOPT", "0") == "1" or not sys.platform.startswith("linux"):
        return False
    try:
        import ctypes
        libc = ctypes.CDLL("libc.so.6")
        m_trim_threshold, m_mmap_threshold = -1, -3       # <malloc.h>
        return bool(libc.mallopt(m_mmap_threshold, 1 << 30) and libc.mallopt(m_trim_threshold, 1 << 30))
    except (OSError, AttributeError):
        return False


def freeze_long_lived():
    """Take what exists now -- above all the chain's Datum objects with their dicts and lists, 50k+ containers that live as long
    as the run -- out of the cyclic garbage collector's sight (gc.freeze).  A full collection walks every tracked container:
    10-20 ms at 50k SSMs, several times per hundred tree samples, for objects that can never become garbage.  Reference
    counting and the collection of what is allocated later are unaffected.  PWGS_NO_GC_FREEZE=1 turns it off."""
    if os.environ.get("PWGS_NO_GC_FREEZE", "0") == "1":
        return False
    import gc
    gc.collect()
    gc.freeze()
    return True


def rm_safely(fn):
    try:
        os.remove(fn)
    except FileNotFoundError:
        pass


class BackupManager(object):
    def __init__(self, filenames):
        self._pairs = [(fn, os.path.realpath(fn) + ".backup") for fn in filenames]

    def save_backup(self):
        for fn, bak in self._pairs:
            shutil.copy2(fn, bak)

    def restore_backup(self):
        for fn, bak in self._pairs:
            shutil.copy2(bak, fn)

    def remove_backup(self):
        for _, bak in self._pairs:
            rm_safely(bak)


class StateManager(object):
    default_last_state_fn = "state.last.pickle"
    default_initial_state_fn = "state.initial.pickle"

    def __init__(self, workdir="."):
        # every chain file is addressed through its directory, never through the process' cwd: several chains can then
        # share one process (multievolve --in-process)
        self.last_state_fn = os.path.join(workdir, self.default_last_state_fn)
        self.initial_state_fn = os.path.join(workdir, self.default_initial_state_fn)

    @staticmethod
    def dumps(state):
        tssb = state.get("tssb")
        plain = tssb is None or os.environ.get("PWGS_PLAIN_PICKLE", "0") == "1"
        return fastpickle.compat_dumps(state) if plain else fastpickle.dumps(tssb, root=state)

    def write_state_bytes(self, blob):
        with open(self.last_state_fn, "wb") as f:
            f.write(blob)

    def _dump(self, state, fn):
        with open(fn, "wb") as f:
            f.write(self.dumps(state))

    def write_state(self, state):
        self._dump(state, self.last_state_fn)

    def write_initial_state(self, state):
        self._dump(state, self.initial_state_fn)

    def load_state(self):
        with open(self.last_state_fn, "rb") as f:
            return pickle.load(f)

    def state_exists(self):
        return os.path.isfile(self.last_state_fn)


class CorruptZipFileError(Exception):
    pass


class TreeWriter(object):
    default_archive_fn = "trees.zip"

    def __init__(self, resume_run=False, workdir="."):
        self.archive_fn = os.path.join(workdir, self.default_archive_fn)
        if resume_run:
            with zipfile.ZipFile(self.archive_fn) as z:
                if z.testzip() is not None:
                    raise CorruptZipFileError("Corrupt zip file: %s" % self.archive_fn)
        else:
            rm_safely(self.archive_fn)

    def _open(self):
        return zipfile.ZipFile(self.archive_fn, "a", compression=zipfile.ZIP_DEFLATED, allowZip64=True, compresslevel=1)

    def add_extra_file(self, filename, data):
        with self._open() as z:
            z.writestr(filename, data)

    def write_trees(self, serialized_trees):
        with self._open() as z:
            for blob, idx, llh in serialized_trees:
                z.writestr("%s_%s_%s" % ("burnin" if idx < 0 else "tree", idx, llh), blob)


class AsyncWriter(object):
    """A chain's checkpoint writes (mcmc_samples.txt, trees.zip, state.last.pickle, backups) on one background thread, in order,
    one checkpoint in flight at a time: deflating ten trees takes as long as a tree sample or two, releases the GIL, and has no
    reason to stall the chain.  What is written is fixed when the job is submitted (the trees and the state are already bytes).
    A signal handler calls stop_and_drain(): no new write starts afterwards and the one in flight is finished, so the files
    are as consistent as with in-line writes (trees first, then the state that refers to them)."""

    def __init__(self):
        self._lock = threading.Lock()
        self._idle = threading.Event()
        self._idle.set()
        self._stopping = False
        self._error = None
        self._jobs = queue.Queue()
        threading.Thread(target=self._run, daemon=True).start()

    def _run(self):
        while True:
            job = self._jobs.get()
            try:
                job()
            except BaseException as e:      # noqa: B902 - handed to the chain's thread by wait()
                self._error = e
            finally:
                self._idle.set()

    def wait(self):
        self._idle.wait()
        if self._error is not None:
            e, self._error = self._error, None
            raise e

    def submit(self, job):
        self.wait()
        with self._lock:
            stopping = self._stopping
            if not stopping:
                self._idle.clear()
                self._jobs.put(job)
        if stopping:                        # the process is about to exit on a signal: never start another write
            threading.Event().wait()

    def stop_and_drain(self):
        with self._lock:
            self._stopping = True
        self._idle.wait()


# ------------------------------------------------------------------ the chain
def chain_device_index(config):
    """The chain's GPU: config["device"] (multievolve --in-process) or PWGS_DEVICE (one process per chain), default 0."""
    dev = config.get("device")
    return int(os.environ.get("PWGS_DEVICE", "0")) if dev is None else int(dev)


def chain_log(config, msg, fd=None):
    logmsg(config.get("log_prefix", "") + msg, fd)


def _attach(state, codes, config):
    """(Re)create the process-local parts of a chain: its generator and its device context."""
    tssb = state["tssb"]
    tssb.rng = np.random.RandomState()
    tssb.backend = GpuBackend(codes, device=chain_device_index(config))
    tssb.data = codes
    for d in codes:
        d.tssb = tssb
    return tssb


def start_new_run(state_manager, backup_manager, safe_to_exit, run_succeeded, config, ssm_file, cnv_file, params_file,
                  burnin_samples, num_samples, mh_itr, mh_std, write_state_every, write_backups_every, rand_seed, tmp_dir,
                  workdir="."):
    state = {}
    seed_fn = os.path.join(workdir, "random_seed.txt")
    if rand_seed is not None:
        state["rand_seed"] = int(rand_seed)
    else:                                       # CLI -> random_seed.txt -> fresh seed (evolve.py:33-53)
        try:
            with open(seed_fn) as f:
                state["rand_seed"] = int(f.read().strip())
        except (OSError, ValueError):
            state["rand_seed"] = int(np.random.randint(2 ** 32, dtype=np.uint64))
    with open(seed_fn, "w") as f:
        f.write("%s\n" % state["rand_seed"])
    state.update(working_dir=os.path.abspath(workdir), ssm_file=ssm_file, cnv_file=cnv_file, tmp_dir=tmp_dir,
                 write_state_every=write_state_every, write_backups_every=write_backups_every)

    codes, n_ssms, n_cnvs, mapping = load_data(ssm_file, cnv_file)
    if not codes:
        chain_log(config, "No SSMs or CNVs provided. Exiting.", sys.stderr)
        return
    ntps = len(codes[0].a)
    state["glist"] = [d.name for d in codes if len(d.name) > 0]
    state.update(burnin=burnin_samples, num_samples=num_samples, dp_alpha=25.0, dp_gamma=1.0, alpha_decay=0.25,
                 mh_burnin=0, mh_itr=mh_itr, mh_std=mh_std)
    state["cd_llh_traces"] = np.zeros((num_samples, 1))
    state["burnin_cd_llh_traces"] = np.zeros((burnin_samples, 1))

    rng = np.random.RandomState(state["rand_seed"])
    backend = GpuBackend(codes, device=chain_device_index(config))
    root = alleles(conc=0.1, ntps=ntps)
    tssb = TSSB(dp_alpha=state["dp_alpha"], dp_gamma=state["dp_gamma"], alpha_decay=state["alpha_decay"], root_node=root,
                data=codes, rng=rng, backend=backend)
    # one child under the (empty) root takes all the data (evolve.py:84-98)
    tssb.root["sticks"] = np.vstack([tssb.root["sticks"], .999999])
    child = root.spawn()
    tssb.root["children"].append({"node": child, "main": boundbeta(rng, 1.0, tssb.alpha_decay * tssb.dp_alpha),
                                  "sticks": np.empty((0, 1)), "children": []})
    for n in range(tssb.num_data):
        root.remove_datum(n)
        child.add_datum(n)
        tssb.assignments[n] = child
    for d in codes:
        d.tssb = tssb
    state["tssb"] = tssb

    writer = TreeWriter(workdir=workdir)
    writer.add_extra_file("cnv_logical_physical_mapping.json", json.dumps(mapping))
    params = {}
    if params_file is not None:
        with open(params_file) as f:
            params = json.load(f)
    writer.add_extra_file("params.json", json.dumps(params))

    state_manager.write_initial_state(state)
    chain_log(config, "Starting MCMC run...")
    state["last_iteration"] = -state["burnin"] - 1
    with open(os.path.join(workdir, "mcmc_samples.txt"), "w") as f:
        f.write("Iteration\tLLH\tTime\n")
    do_mcmc(state_manager, backup_manager, safe_to_exit, run_succeeded, config, state, writer, codes, n_ssms, n_cnvs, ntps, tmp_dir,
            workdir)


def resume_existing_run(state_manager, backup_manager, safe_to_exit, run_succeeded, config, workdir="."):
    # the reference chdir()s into state["working_dir"] here (evolve.py:262); that is the directory the state was just read
    # from, and the one every path below is built on
    try:
        state = state_manager.load_state()
        writer = TreeWriter(resume_run=True, workdir=workdir)
    except Exception:
        chain_log(config, "Restoring state failed:", sys.stderr)
        traceback.print_exc()
        chain_log(config, "Restoring from backup and trying again.", sys.stderr)
        backup_manager.restore_backup()
        state = state_manager.load_state()
        writer = TreeWriter(resume_run=True, workdir=workdir)
    state["working_dir"] = os.path.abspath(workdir)
    tssb = state["tssb"]
    codes = tssb.data                           # the pickled data carry the assignments' bookkeeping (datum.node)
    n_ssms = sum(1 for d in codes if d.id[0] == "s")
    _attach(state, codes, config)
    tssb.rng.set_state(state["rand_state"])
    do_mcmc(state_manager, backup_manager, safe_to_exit, run_succeeded, config, state, writer, codes, n_ssms, len(codes) - n_ssms,
            len(codes[0].a), state["tmp_dir"], workdir)


def do_mcmc(state_manager, backup_manager, safe_to_exit, run_succeeded, config, state, tree_writer, codes, n_ssms, n_cnvs, ntps,
            tmp_dir_parent, workdir="."):
    samples_fn = os.path.join(workdir, "mcmc_samples.txt")
    plain_pickle = os.environ.get("PWGS_PLAIN_PICKLE", "0") == "1"
    writer = None if os.environ.get("PWGS_SYNC_WRITES", "0") == "1" else AsyncWriter()
    config["writer"] = writer
    on_status = config.get("on_status")           # multievolve --in-process reads the progress here instead of from a pipe
    start_iter = state["last_iteration"] + 1
    unwritten, sample_times = [], []
    t_last = time.time()
    config["tmp_dir"] = tempfile.mkdtemp(prefix="pwgsdataexchange.", dir=tmp_dir_parent)
    tssb = state["tssb"]
    tune_allocator()
    freeze_long_lived()

    for iteration in range(start_iter, state["num_samples"]):
        sys.stdout.flush()
        safe_to_exit.set()
        status = OrderedDict(iteration=iteration, trees_sampled=state["burnin"] + iteration,
                             total_trees=state["burnin"] + state["num_samples"])

        tssb.resample_assignments()
        tssb.cull_tree()
        _, nodes = tssb.get_mixture()
        for i, nd in enumerate(nodes):
            nd.id = i
        set_node_height(tssb)
        set_path_from_root_to_node(tssb)
        # datum.node (util2.py:111-115): set for every datum once; afterwards the assignment sweep keeps it in step with
        # every move it makes (sampler.py), and nothing else moves data.  The full pass costs ~10 ms at 50k SSMs.
        if iteration == start_iter or os.environ.get("PWGS_CHECK_INVARIANTS") == "1":
            if iteration != start_iter:
                assert all(d.node is tssb.assignments[n] for n, d in enumerate(tssb.data)), "datum.node out of step"
            map_datum_to_node(tssb)

        state["mh_acc"] = metropolis(tssb, state["mh_itr"], state["mh_std"], state["mh_burnin"], n_ssms, n_cnvs, state["ssm_file"],
                                     state["cnv_file"], state["rand_seed"], ntps, config["tmp_dir"],
                                     stream=iteration + state["burnin"])
        acc = float(state["mh_acc"])
        if acc < 0.08 and state["mh_std"] < 10000:
            state["mh_std"] = state["mh_std"] * 2.0
            chain_log(config, "Shrinking MH proposals. Now %f" % state["mh_std"])
        if 0.5 < acc < 0.99:
            state["mh_std"] = state["mh_std"] / 2.0
            chain_log(config, "Growing MH proposals. Now %f" % state["mh_std"])

        tssb.resample_sticks()
        tssb.resample_stick_orders()
        tssb.resample_hypers(dp_alpha=True, alpha_decay=True, dp_gamma=True)

        last_llh = tssb.complete_data_log_likelihood()
        status["llh"] = last_llh
        if iteration >= 0:
            state["cd_llh_traces"][iteration] = last_llh
            status["nodes"] = len(tssb.get_nodes())
            status["mh_acc"] = state["mh_acc"]
            status["dp_alpha"] = tssb.dp_alpha
            status["dp_gamma"] = tssb.dp_gamma
            status["alpha_decay"] = tssb.alpha_decay
        else:
            state["burnin_cd_llh_traces"][iteration + state["burnin"]] = last_llh
        if on_status is not None:
            on_status(status)
        else:
            logmsg(" ".join("%s=%s" % kv for kv in status.items()))

        # protocol-2 pickle of the tree (util2.py:250-262), assembled from cached per-datum fragments; PWGS_PLAIN_PICKLE=1 = pickle.dumps
        unwritten.append((fastpickle.compat_dumps(tssb) if plain_pickle else fastpickle.dumps(tssb), iteration, last_llh))
        state["rand_state"] = tssb.rng.get_state()
        state["last_iteration"] = iteration
        if len([c for c in tssb.root["children"] if c["node"].has_data()]) > 1:
            chain_log(config, "Polyclonal tree detected with %s clones." % len(tssb.root["children"]))

        now = time.time()
        sample_times.append(now - t_last)
        t_last = now

        safe_to_exit.clear()                    # no exit in the middle of a zip / pickle write (evolve.py:247-250)
        write_backup = iteration % state["write_backups_every"] == 0 and iteration != start_iter
        write_state = iteration % state["write_state_every"] == 0
        if write_backup or write_state or iteration == state["num_samples"] - 1:
            rows = "\n".join("%s\t%s\t%s" % (it, llh, dt) for (_, it, llh), dt in zip(unwritten, sample_times)) + "\n"
            trees, state_blob = unwritten, state_manager.dumps(state)     # fixed now; the chain goes on while they are written

            def checkpoint(rows=rows, trees=trees, state_blob=state_blob, write_backup=write_backup):
                with open(samples_fn, "a") as f:
                    f.write(rows)
                tree_writer.write_trees(trees)
                state_manager.write_state_bytes(state_blob)
                if write_backup:
                    backup_manager.save_backup()
            if writer is None:
                checkpoint()
            else:
                writer.submit(checkpoint)
            unwritten, sample_times = [], []

    if writer is not None:
        writer.wait()
    backup_manager.remove_backup()
    safe_to_exit.set()
    run_succeeded.set()


# ------------------------------------------------------------------ CLI (evolve.py:287-497)
def create_argparser(full=True):
    p = argparse.ArgumentParser(description="Run PhyloWGS to infer subclonal composition from SSMs and CNVs",
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("-O", "--output-dir", dest="output_dir", help="Path to output directory")
    if full:
        p.add_argument("-b", "--write-backups-every", dest="write_backups_every", default=100, type=int,
                       help="Number of iterations to go between writing backups of program state")
        p.add_argument("-S", "--write-state-every", dest="write_state_every", default=10, type=int,
                       help="Number of iterations between writing program state to disk")
        p.add_argument("-B", "--burnin-samples", dest="burnin_samples", default=1000, type=int, help="Number of burnin samples")
        p.add_argument("-s", "--mcmc-samples", dest="mcmc_samples", default=2500, type=int, help="Number of MCMC samples")
        p.add_argument("-i", "--mh-iterations", dest="mh_iterations", default=5000, type=int,
                       help="Number of Metropolis-Hastings iterations")
        p.add_argument("-r", "--random-seed", dest="random_seed", type=int, help="Random seed for initializing MCMC sampler")
        p.add_argument("-t", "--tmp-dir", dest="tmp_dir", help="Path to directory for temporary files")
        p.add_argument("-p", "--params", dest="params_file", help="JSON file listing run parameters, generated by the parser")
        p.add_argument("ssm_file", help="File listing SSMs. For proper format, see README.md.")
        p.add_argument("cnv_file", help="File listing CNVs. For proper format, see README.md.")
    return p


def run(safe_to_exit, run_succeeded, config, argv=None):
    install_pickle_aliases()
    known, _ = create_argparser(full=False).parse_known_args(argv)
    orig_dir = os.getcwd()
    workdir = orig_dir                            # the reference chdir()s here (evolve.py:341-344); we carry the path instead
    if known.output_dir is not None:
        workdir = os.path.abspath(known.output_dir)
        os.makedirs(workdir, exist_ok=True)
    state_manager = StateManager(workdir)
    backup_manager = BackupManager([state_manager.last_state_fn, os.path.join(workdir, TreeWriter.default_archive_fn)])
    if state_manager.state_exists():
        chain_log(config, "Resuming existing run. Ignoring command-line parameters (except --output-dir).")
        resume_existing_run(state_manager, backup_manager, safe_to_exit, run_succeeded, config, workdir)
        return
    args = create_argparser().parse_args(argv)
    paths = {}
    for key in ("ssm_file", "cnv_file", "params_file", "tmp_dir"):
        v = getattr(args, key)
        paths[key] = os.path.normpath(os.path.join(orig_dir, v)) if v is not None else None
    for key in ("ssm_file", "cnv_file"):
        try:
            open(paths[key]).close()
        except OSError as e:
            sys.stderr.write(str(e) + "\n")
            sys.exit(1)
    start_new_run(state_manager, backup_manager, safe_to_exit, run_succeeded, config, paths["ssm_file"], paths["cnv_file"],
                  paths["params_file"], burnin_samples=args.burnin_samples, num_samples=args.mcmc_samples,
                  mh_itr=args.mh_iterations, mh_std=100, write_state_every=args.write_state_every,
                  write_backups_every=args.write_backups_every, rand_seed=args.random_seed, tmp_dir=paths["tmp_dir"],
                  workdir=workdir)


def remove_tmp_files(tmp_dir):
    if tmp_dir is None:
        return
    for fn in get_c_fnames(tmp_dir):
        rm_safely(fn)
    try:
        os.rmdir(tmp_dir)
    except OSError:
        pass


def main(argv=None):
    if argv is None:
        argv = sys.argv[1:]
    if "-h" in argv or "--help" in argv:
        create_argparser().print_help()
        sys.exit()
    safe_to_exit, run_succeeded = threading.Event(), threading.Event()
    config = {"tmp_dir": None}

    def on_signal(signo, _frame):
        logmsg("Signal %s received." % signo, sys.stderr)
        safe_to_exit.wait()
        if config.get("writer") is not None:
            config["writer"].stop_and_drain()
        remove_tmp_files(config["tmp_dir"])
        logmsg("Exiting now.")
        sys.exit(3)

    signal.signal(signal.SIGTERM, on_signal)
    signal.signal(signal.SIGINT, on_signal)
    worker = threading.Thread(target=run, args=(safe_to_exit, run_succeeded, config, argv), daemon=True)
    worker.start()
    while worker.is_alive():
        worker.join(10)
    remove_tmp_files(config["tmp_dir"])
    if run_succeeded.is_set():
        logmsg("Run succeeded.")
        sys.exit(0)
    logmsg("Run failed.")
    sys.exit(1)


if __name__ == "__main__":
    main()
