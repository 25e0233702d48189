"""Experiment: snapshot-pickling of the TSSB in a forked child (copy-on-write) while the chain goes on.  Measures what the parent
pays (fork latency, reading the bytes back) against pickling in-line, inside a process with a live CUDA context."""
import os
import pickle
import sys
import tempfile
import threading
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402
from phylowgs_b200.host import evolve as E  # noqa: E402
from phylowgs_b200.host.backend import GpuBackend  # noqa: E402
from phylowgs_b200.host.model import TSSB, alleles, boundbeta  # noqa: E402
from phylowgs_b200.host.params import metropolis  # noqa: E402


def build(s):
    with tempfile.TemporaryDirectory() as wd:
        data = O.Data(s["a"], s["d"], s["mu_r"], s["mu_v"], s["n_ssm"], s["cnv_ptr"], s["cnv_idx"], s["cnv_c1"], s["cnv_c2"])
        ssm, cnv = os.path.join(wd, "ssm_data.txt"), os.path.join(wd, "cnv_data.txt")
        O.write_ssm_cnv(data, ssm, cnv)
        codes, n_ssms, n_cnvs, _ = E.load_data(ssm, cnv)
    E.install_pickle_aliases()
    rng = np.random.RandomState(1)
    backend = GpuBackend(codes, device=0)
    root = alleles(conc=0.1, ntps=len(codes[0].a))
    tssb = TSSB(dp_alpha=25.0, dp_gamma=1.0, alpha_decay=0.25, root_node=root, data=codes, rng=rng, backend=backend)
    tssb.root["sticks"] = np.vstack([tssb.root["sticks"], .999999])
    child = root.spawn()
    tssb.root["children"].append({"node": child, "main": boundbeta(rng, 1.0, tssb.alpha_decay * tssb.dp_alpha), "sticks": np.empty((0, 1)), "children": []})
    for n in range(tssb.num_data):
        root.remove_datum(n); child.add_datum(n); tssb.assignments[n] = child
    for d in codes:
        d.tssb = tssb
    return tssb, n_ssms, n_cnvs


def main():
    s, _ = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "config3")
    tssb, n_ssms, n_cnvs = build(s)
    for it in range(2):
        tssb.resample_assignments(); tssb.cull_tree()
        metropolis(tssb, 5000, 100.0, 0, n_ssms, n_cnvs, "", "", 1, 1, ".", stream=it)
    t0 = time.perf_counter(); blob = pickle.dumps(tssb, protocol=2); t1 = time.perf_counter()
    print("in-line pickle: %.1f ms, %d bytes" % (1e3 * (t1 - t0), len(blob)), flush=True)
    for rep in range(3):
        r, w = os.pipe()
        t0 = time.perf_counter()
        pid = os.fork()
        if pid == 0:
            try:
                os.close(r)
                b = pickle.dumps(tssb, protocol=2)
                with os.fdopen(w, "wb") as f:
                    f.write(b)
            finally:
                os._exit(0)
        t1 = time.perf_counter()
        os.close(w)
        # the chain goes on: one more iteration while the child pickles
        tssb.resample_assignments(); tssb.cull_tree()
        metropolis(tssb, 5000, 100.0, 0, n_ssms, n_cnvs, "", "", 1, 1, ".", stream=10 + rep)
        t2 = time.perf_counter()
        with os.fdopen(r, "rb") as f:
            got = f.read()
        os.waitpid(pid, 0)
        t3 = time.perf_counter()
        print("fork returned after %.1f ms; next sweep+MH %.1f ms; bytes back after another %.1f ms; equal to a snapshot taken in-line at fork time: %s (%d bytes)"
              % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), got == blob if rep == 0 else "n/a", len(got)), flush=True)
    tssb.backend.close()


th = threading.Thread(target=main)
th.start(); th.join()
