"""End-to-end statistical comparison of chains (SURVEY.md 8c last row).  TEST / EVIDENCE TOOLING.

    python scripts/stat_parity.py ref  --set config2 --seed 1 -O <dir>      one REFERENCE chain (build container only)
    python scripts/stat_parity.py gpu  --set config2 --seed 1 -O <dir>      one chain of this repo on cuda:0
    python scripts/stat_parity.py summarise <dir> [<dir> ...] -o out.json   what the comparison needs from finished chains
    python scripts/stat_parity.py compare ref.json gpu.json                 the tolerances of SURVEY 8c, as a table + exit code

The reference chain is the reference's OWN evolve.py / tssb.py / data.py / params.py (executed under Python 3 by oracle/ref_py3.py,
declared mechanical patches only) driving the UNMODIFIED mh.cpp (oracle/_ref/mh.o, g++ -O3, GSL stand-in oracle/gsl_shim)
through its own text files — no code of this repo computes anything in it.  One declared acceleration, used only for sets
whose data carry no CNV links (`accel` below): util2.set_node_height / set_path_from_root_to_node / map_datum_to_node, which
data.py:33-38 re-runs on EVERY likelihood call (O(N) each, 140 s per tree sample at 10k SSMs), are no-ops while a sweep's
likelihoods are evaluated.  For a datum without CNV links the likelihood reads none of what they maintain (data.py:59-61), and
evolve.py:180-182 runs its own copies before every metropolis() call, so every number of the chain is unchanged.

Quantities (those the reference itself records): complete-data LLH trace (mcmc_samples.txt, evolve.py:213-217), nodes per sample
(status line `nodes=`, evolve.py:219), populations that hold data, MH acceptance ratio (`mh_acc=`), per-SSM posterior mean phi
(node.params of the SSM's node, result_generator.py:57) over the post-burn-in trees.
"""
import argparse
import json
import os
import pickle
import re
import sys
import threading
import time
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

SETS = {
    # BASELINE.json configs[0]: bundled example, --burnin-samples 100 --mcmc-samples 250
    "config1": dict(ssm=os.path.join(GOLDEN, "ssm_data.txt"), cnv=os.path.join(GOLDEN, "cnv_data.txt"), burnin=100, samples=250,
                    mh_iters=5000, accel=False),
    # a small phylowgs_sims file: the unaccelerated reference is affordable
    "sim150": dict(ssm=os.path.join(GOLDEN, "emptysim.4.100.50.ssm.txt"), cnv=None, burnin=100, samples=200, mh_iters=5000, accel=False),
    # BASELINE.json configs[1]: emptysim.6.300.2000 (10k SSMs); shortened from 1000 / 2500 to what the CPU side finishes here
    "config2": dict(ssm=os.path.join(GOLDEN, "emptysim.6.300.2000.ssm.txt"), cnv=None, burnin=150, samples=250, mh_iters=5000, accel=True),
}


def empty_cnv_file(outdir):
    fn = os.path.join(outdir, "empty_cnv.txt")
    open(fn, "w").close()
    return fn


def run_ref(args):
    from oracle import ref_py3
    cfg = SETS[args.set]
    os.makedirs(args.output, exist_ok=True)
    cnv = cfg["cnv"] or empty_cnv_file(os.path.abspath(args.output))
    R = ref_py3.load()
    if cfg["accel"]:
        noop = lambda tssb: None          # noqa: E731  (see the module docstring: result-neutral without CNV links)
        R.util2.set_node_height = noop
        R.util2.set_path_from_root_to_node = noop
        R.util2.map_datum_to_node = noop
        # params.metropolis (params.py:46,48) calls the same two through util2: evolve.py:180-182 has just run its own copies
    os.chdir(args.output)
    sys.stdout.flush()
    saved_fd = os.dup(1)                  # at the descriptor level: util2.logmsg binds sys.stdout when it is defined
    log = open("chain.log", "w")
    os.dup2(log.fileno(), 1)
    t0 = time.time()
    E = R.evolve
    sm = R.util2.StateManager()
    bm = R.util2.BackupManager([R.util2.StateManager.default_last_state_fn, R.util2.TreeWriter.default_archive_fn])
    E.start_new_run(sm, bm, threading.Event(), threading.Event(), {"tmp_dir": None}, cfg["ssm"], cnv, None,
                    burnin_samples=args.burnin or cfg["burnin"], num_samples=args.samples or cfg["samples"],
                    mh_itr=args.mh_iters or cfg["mh_iters"], mh_std=100, write_state_every=10, write_backups_every=1000000,
                    rand_seed=args.seed, tmp_dir=None)
    sys.stdout.flush()
    os.dup2(saved_fd, 1)
    log.close()
    with open("wall_seconds.txt", "w") as f:
        f.write("%f\n" % (time.time() - t0))


def run_gpu(args):
    from phylowgs_b200.host import evolve as E
    cfg = SETS[args.set]
    os.makedirs(args.output, exist_ok=True)
    cnv = cfg["cnv"] or empty_cnv_file(os.path.abspath(args.output))
    sys.stdout.flush()
    saved_fd = os.dup(1)
    log = open(os.path.join(args.output, "chain.log"), "w")
    os.dup2(log.fileno(), 1)
    t0 = time.time()
    safe, ok = threading.Event(), threading.Event()
    config = {"tmp_dir": None}
    E.run(safe, ok, config, ["-O", args.output, "-B", str(args.burnin or cfg["burnin"]), "-s", str(args.samples or cfg["samples"]),
                             "-i", str(args.mh_iters or cfg["mh_iters"]), "-r", str(args.seed), "-S", "50", cfg["ssm"], cnv])
    E.remove_tmp_files(config["tmp_dir"])
    sys.stdout.flush()
    os.dup2(saved_fd, 1)
    log.close()
    assert ok.is_set(), "chain failed"
    with open(os.path.join(args.output, "wall_seconds.txt"), "w") as f:
        f.write("%f\n" % (time.time() - t0))


def summarise_chain(d):
    """One finished chain directory -> dict.  Trees are unpickled with this repo's model classes under the reference's module
    names (both sides pickle the same attribute layout)."""
    from phylowgs_b200.host import evolve as E
    E.install_pickle_aliases()
    rows = [r.split("\t") for r in open(os.path.join(d, "mcmc_samples.txt")).read().strip().split("\n")[1:]]
    it = np.array([int(r[0]) for r in rows])
    llh = np.array([float(r[1]) for r in rows])
    secs = np.array([float(r[2]) for r in rows])
    status = {}
    for line in open(os.path.join(d, "chain.log")):
        m = re.search(r"iteration=(-?\d+) .*? llh=(\S+)(?: nodes=(\d+) mh_acc=(\S+))?", line)
        if m and m.group(3):
            status[int(m.group(1))] = (int(m.group(3)), float(m.group(4)))
    post = sorted(k for k in status if k >= 0)
    phi_sum, n_trees, pops, big_pops = None, 0, [], []
    with zipfile.ZipFile(os.path.join(d, "trees.zip")) as z:
        for name in z.namelist():
            if not name.startswith("tree_"):
                continue
            t = pickle.loads(z.read(name))
            n_ssm = sum(1 for dat in t.data if dat.id[0] == "s")
            phis = np.array([np.asarray(t.assignments[n].params, dtype=float).reshape(-1) for n in range(len(t.data)) if t.data[n].id[0] == "s"])
            phi_sum = phis if phi_sum is None else phi_sum + phis
            n_trees += 1
            counts = [len(nd.data) for nd in t.get_nodes()]
            pops.append(sum(1 for c in counts if c > 0))
            big_pops.append(sum(1 for c in counts if c >= max(2, 0.01 * n_ssm)))
    wall = float(open(os.path.join(d, "wall_seconds.txt")).read()) if os.path.exists(os.path.join(d, "wall_seconds.txt")) else None
    return dict(dir=os.path.basename(os.path.normpath(d)), n_post=int(np.sum(it >= 0)), mean_llh=float(np.mean(llh[it >= 0])),
                sd_llh=float(np.std(llh[it >= 0])), last_llh=float(llh[-1]), llh_trace=[float(x) for x in llh],
                nodes_mean=float(np.mean([status[k][0] for k in post])), mh_acc_mean=float(np.mean([status[k][1] for k in post])),
                mh_acc_trace=[status[k][1] for k in post], nodes_trace=[status[k][0] for k in post],
                pops_mode=int(np.bincount(pops).argmax()), pops_mean=float(np.mean(pops)),
                big_pops_mode=int(np.bincount(big_pops).argmax()), big_pops_mean=float(np.mean(big_pops)),
                mean_phi=(phi_sum / n_trees).tolist(), n_trees=n_trees, seconds_per_sample=float(np.median(secs)), wall_seconds=wall)


def summarise(args):
    out = dict(chains=[summarise_chain(d) for d in args.dirs])
    # per-SSM posterior mean phi: the average over the chains is what the comparison uses; kept once, rounded, to keep the file small
    out["mean_phi"] = np.round(np.mean([np.array(c["mean_phi"]) for c in out["chains"]], axis=0), 6).tolist()
    for c in out["chains"]:
        c["llh_trace"] = [round(x, 3) for x in c["llh_trace"]]
        phi = c.pop("mean_phi")
        if len(phi) <= 500:                                  # small sets: per-chain means too (chains stuck in different modes)
            c["mean_phi"] = np.round(np.array(phi), 5).tolist()
    with open(args.out, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    for c in out["chains"]:
        print("%-24s post=%d mean_llh=%.2f sd=%.2f nodes=%.1f pops(mode)=%d big_pops(mode)=%d mh_acc=%.4f s/sample=%.3f"
              % (c["dir"], c["n_post"], c["mean_llh"], c["sd_llh"], c["nodes_mean"], c["pops_mode"], c["big_pops_mode"], c["mh_acc_mean"],
                 c["seconds_per_sample"]))


def compare_summaries(ref, gpu):
    """SURVEY 8c's tolerances.  -> (rows, ok)."""
    r, g = ref["chains"], gpu["chains"]
    rows, ok = [], True
    r_llh, g_llh = np.array([c["mean_llh"] for c in r]), np.array([c["mean_llh"] for c in g])
    sd = max(float(np.std(r_llh, ddof=1)) if len(r) > 1 else 0.0, float(np.mean([c["sd_llh"] for c in r])) / np.sqrt(max(r[0]["n_post"], 1)), 1e-9)
    gap = abs(float(np.mean(g_llh)) - float(np.mean(r_llh)))
    rows.append(("post-burn-in mean LLH", "%.2f" % np.mean(r_llh), "%.2f" % np.mean(g_llh), "gap %.2f <= 3 x between-seed SD %.2f" % (gap, sd), gap <= 3 * sd))
    r_pop = int(np.bincount([c["big_pops_mode"] for c in r]).argmax())
    g_pop = int(np.bincount([c["big_pops_mode"] for c in g]).argmax())
    rows.append(("modal cluster count (populations holding >= 1 % of the SSMs)", str(r_pop), str(g_pop), "equal", r_pop == g_pop))
    r_phi, g_phi = np.array(ref["mean_phi"]), np.array(gpu["mean_phi"])
    rmse = float(np.sqrt(np.mean((r_phi - g_phi) ** 2)))
    rows.append(("per-SSM posterior mean phi", "-", "-", "RMSE %.4f <= 0.02" % rmse, rmse <= 0.02))
    r_acc, g_acc = float(np.mean([c["mh_acc_mean"] for c in r])), float(np.mean([c["mh_acc_mean"] for c in g]))
    rows.append(("MH acceptance ratio (mean over samples)", "%.4f" % r_acc, "%.4f" % g_acc, "|diff| %.4f <= 0.05" % abs(r_acc - g_acc), abs(r_acc - g_acc) <= 0.05))
    rows.append(("nodes per sample (mean; reported, no tolerance)", "%.1f" % np.mean([c["nodes_mean"] for c in r]), "%.1f" % np.mean([c["nodes_mean"] for c in g]), "-", True))
    rows.append(("seconds per tree sample (median)", "%.3f" % np.median([c["seconds_per_sample"] for c in r]),
                 "%.4f" % np.median([c["seconds_per_sample"] for c in g]), "-", True))
    # Chains that do not mix (config 1: MH acceptance 2e-5 on both sides, every chain frozen in the mode its burn-in found, modes
    # thousands of log-likelihood units apart) make the pooled mean phi a statement about HOW OFTEN each mode was found, not about
    # the likelihood.  So: chains are grouped into modes by their mean LLH, the occupancy is reported, and phi is compared mode by
    # mode where both sides have chains (same tolerance); the pooled row then counts only if the chains mix.
    if all("mean_phi" in c for c in r + g):
        width = 10.0 * max(float(np.mean([c["sd_llh"] for c in r + g])), 1.0)
        centres = []
        for v in sorted(np.concatenate([r_llh, g_llh])):
            if not centres or v - centres[-1][-1] > width:
                centres.append([v])
            else:
                centres[-1].append(v)
        if len(centres) > 1:
            modes = [(min(c) - 1e-6, max(c) + 1e-6) for c in centres]
            worst, shared, occ = 0.0, 0, []
            for lo, hi in modes:
                rr = [np.array(c["mean_phi"]) for c in r if lo <= c["mean_llh"] <= hi]
                gg = [np.array(c["mean_phi"]) for c in g if lo <= c["mean_llh"] <= hi]
                occ.append("%.0f: %d / %d" % (0.5 * (lo + hi), len(rr), len(gg)))
                if rr and gg:
                    shared += 1
                    worst = max(worst, float(np.sqrt(np.mean((np.mean(rr, axis=0) - np.mean(gg, axis=0)) ** 2))))
            rows[2] = (rows[2][0] + " (pooled over chains that do not mix: reported, no tolerance)", "-", "-", "RMSE %.4f" % rmse, True)
            rows.insert(3, ("per-SSM mean phi, mode by mode (%d LLH modes, %d found by both sides)" % (len(modes), shared), "-", "-",
                            "worst RMSE %.4f <= 0.02" % worst, shared > 0 and worst <= 0.02))
            rows.insert(4, ("chains per LLH mode (mode: reference / this repo; reported)", "-", "-", "; ".join(occ), True))
    for row in rows:
        ok = ok and row[4]
    return rows, ok


def compare(args):
    ref, gpu = json.load(open(args.ref)), json.load(open(args.gpu))
    rows, ok = compare_summaries(ref, gpu)
    print("| quantity | reference chains (%d) | this repo (%d) | tolerance | ok |" % (len(ref["chains"]), len(gpu["chains"])))
    print("|---|---|---|---|---|")
    for q, a, b, tol, good in rows:
        print("| %s | %s | %s | %s | %s |" % (q, a, b, tol, "yes" if good else "NO"))
    sys.exit(0 if ok else 1)


def main():
    p = argparse.ArgumentParser()
    sub = p.add_subparsers(dest="cmd", required=True)
    for name in ("ref", "gpu"):
        s = sub.add_parser(name)
        s.add_argument("--set", required=True, choices=sorted(SETS))
        s.add_argument("--seed", type=int, required=True)
        s.add_argument("-O", "--output", required=True)
        s.add_argument("--burnin", type=int, default=0)
        s.add_argument("--samples", type=int, default=0)
        s.add_argument("--mh-iters", dest="mh_iters", type=int, default=0)
    s = sub.add_parser("summarise")
    s.add_argument("dirs", nargs="+")
    s.add_argument("-o", "--out", required=True)
    s = sub.add_parser("compare")
    s.add_argument("ref")
    s.add_argument("gpu")
    args = p.parse_args()
    dict(ref=run_ref, gpu=run_gpu, summarise=summarise, compare=compare)[args.cmd](args)


if __name__ == "__main__":
    main()
