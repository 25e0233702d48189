#!/usr/bin/env python
"""bench.py — MH iterations/s of the PhyloWGS hot path on B200 (metric of BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload config1|config2|config3|config4|config5:<n_ssm>[:<T>[:nocnv]]]

One "step" = one `params.metropolis()`-sized pass of the hot path: MH_ITR = 5000 Metropolis-Hastings iterations
(evolve.py:318 default) over the synthetic WGS tumour of BASELINE.json configs[2] (50k SSMs, 300 CNVs, 8 true
subclones, T = 1; phylowgs_b200/synth.py, seed 1003) at the fixed truth tree + 2 empty nodes, mh_std = 100
(SURVEY.md 8d "Bench state").  N > 1 (torchrun): one independent chain per GPU, no data-path collective
(multievolve.py chains share nothing) => weak scaling; value = iterations of all chains / max-over-ranks time.

  value     device-resident: data, tree and state already in HBM; K back-to-back kernel launches between two
            CUDA events on the library's stream (an L2 flush is enqueued between launches).
  e2e       the same K steps through the C-ABI calls a host makes per tree sample (pwgs_set_tree with HOST
            arrays -> pwgs_mh_run -> HOST phi/pi/ratio), wall-clocked, copies inside the timed region.
  roofline  the MH kernel is FP64-pipe-bound (its working set, ~2.6 MB, is on-chip for all 5000 iterations), so
            `bound` is "fp64": achieved = algorithmic FP64-pipe thread-instructions per launch / launch time,
            peak = DFMA issue rate measured live on the same GPU (pwgs_fp64_peak); the HBM view
            (algorithmic bytes per iteration x it/s vs MEASURED_PEAKS.json hbm_gbs) is reported beside it.
  cpu_baseline / --impl reference   the UNMODIFIED reference mh.cpp+util.cpp (oracle/_ref/mh.o, g++ -O3, GSL
            replaced by oracle/gsl_shim because libgsl is not installable here) on the same inputs written in
            the reference's text formats; single-threaded => 1 core per chain.
Nothing here reads /root/reference at run time.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MH_ITR = 5000          # evolve.py:318 (--mh-iterations default)
MH_STD = 100.0         # evolve.py:387
METRIC = "mh_iterations_per_s"
UNIT = "it/s"

# FP64-pipe instruction model of ONE proposal-likelihood pass (DFMA/DADD/DMUL/DSETP/DMNMX thread-instructions; conversions and
# MUFU run on the XU pipe), read off the cubin: profiles/r02_mh_kernel_sass_histogram.txt (scripts/sass_histogram.py) lists the
# loops of mh_kernel<true> with their opcode counts -- the plain loop has 88 FP64 instructions per round of 4 proposals (22 per
# term), the CNV state body 28 per proposal beside its two DFMAs per entry.  (Round 1's fdlibm-style logarithm: 16 per log, 36
# per plain term.)
FP64_LOG = 9                      # pwgs_log_pos_n (table-driven): r, k ln2 + log c, r^2, four Horner steps, r + r^2 P, final add
FP64_EXP = 10                     # pwgs_exp_nonpos_n: clamp, k, k - magic, two-part reduction, r^2, three polynomial steps, scale
FP64_PLAIN_TERM = 2 + 2 * FP64_LOG + 2                     # mu (one fma), 1-mu ; two logs ; two fmas into the sum           = 22
FP64_CNV_STATE = 10 + 2 * FP64_LOG                         # nr+nv, mixture mean (3 + division 3), test, 1-mu, two logs, combine = 28
FP64_CNV_PER_ENTRY = 2                                     # two DFMAs per sparse (node, state) entry
FP64_LSE_EXTRA_STATE = 4 + FP64_EXP                        # logsumexp: max, difference, ONE exp, 1 + e (or the running-sum fma) = 14
FP64_LSE_FINAL = FP64_LOG + 1                              # log of the sum (only when an SSM has more than one merged state)


def workload(name):
    """-> (dict of arrays: data + the fixed tree / assignment of the bench state, description)."""
    from phylowgs_b200 import synth
    if name == "config3":
        return synth.config(3), "synthetic WGS tumour: 50k SSMs, 300 CNVs, 8 true subclones, T=1 (BASELINE.json configs[2])"
    if name == "config4":
        return synth.config(4), "synthetic multi-region: 20k SSMs x 5 samples (BASELINE.json configs[3])"
    if name == "config1":
        return synth.from_files(os.path.join(ROOT, "tests", "golden", "ssm_data.txt"), os.path.join(ROOT, "tests", "golden", "cnv_data.txt"), 2), \
            "the reference's bundled example: 11 SSMs + 1 CNV x 5 samples (BASELINE.json configs[0]), two-population chain tree"
    if name == "config2":
        return synth.from_files(os.path.join(ROOT, "tests", "golden", "emptysim.6.300.2000.ssm.txt"), None, 6), \
            "phylowgs_sims emptysim.6.300.2000: 10k SSMs, no CNVs, T=1 (BASELINE.json configs[1]), six-population chain tree"
    if name.startswith("config5"):                     # config5:<n_ssm>[:<T>[:nocnv]]  (BASELINE.json configs[4]: 1k-500k SSMs)
        tok = name.split(":")
        n = int(tok[1]) if len(tok) > 1 else 100000
        T = int(tok[2]) if len(tok) > 2 else 1
        nocnv = len(tok) > 3 and tok[3] == "nocnv"
        s = synth.make_synthetic(n, 0 if nocnv else max(1, n * 3 // 500), T, 8, seed=1005, frac_cnv=0.0 if nocnv else 0.3, depth=60)
        return s, "scaling sweep: %d SSMs, %s, T=%d (BASELINE.json configs[4])" % (n, "no CNVs" if nocnv else "30 %% of them in %d CNVs" % max(1, n * 3 // 500), T)
    raise SystemExit("unknown workload " + name)


# Round 1's instruction model (fdlibm-form logarithm, 16 FP64 instructions; 36 per plain term ...): kept so that rounds can be
# compared on ONE yardstick of work per iteration -- a cheaper logarithm lowers the executed-instruction fraction while the kernel
# gets faster, see DESIGN.md section 6
R1_MODEL = dict(log=16, plain=36, cnv_state=42, per_entry=2, lse_extra=24, lse_final=17)


def algorithmic_work_r1(s, stats):
    D, T = s["a"].shape
    M = int(np.count_nonzero(np.diff(s["cnv_ptr"]) > 0))
    n1, n2, n3, es = stats
    m = R1_MODEL
    cnv = m["cnv_state"] * (n1 + 2 * n2 + 3 * n3) + m["per_entry"] * es + m["lse_extra"] * (n2 + 2 * n3) + m["lse_final"] * (n2 + n3) + M
    return float(T * (m["plain"] * (D - M) + cnv))


def algorithmic_work(s, K, stats=None):
    """Per MH iteration = ONE proposal-likelihood pass (current-state LLH cached): (bytes, fp64 instructions).
    bytes: the records one pass streams; fp64: the instruction model above applied to the sparse CNV work list
    (stats = counts of CNV-affected SSMs with 1/2/3 merged states and sum of entries x states, from the library)."""
    D, T = s["a"].shape
    M = int(np.count_nonzero(np.diff(s["cnv_ptr"]) > 0))
    by = (D - M) * (20 + 8 * T) + M * (12 + 8 * T + 8 * K)
    if stats is None:
        fl = None
    else:
        n1, n2, n3, es = stats
        cnv = FP64_CNV_STATE * (n1 + 2 * n2 + 3 * n3) + FP64_CNV_PER_ENTRY * es + FP64_LSE_EXTRA_STATE * (n2 + 2 * n3) \
            + FP64_LSE_FINAL * (n2 + n3) + M
        fl = float(T * (FP64_PLAIN_TERM * (D - M) + cnv))
    return by, fl, D, M, T


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML from a host thread DURING the timed region (the recipe's clocks line;
    NVML instead of spawning nvidia-smi so that starting the sampler does not itself disturb the timed kernels)."""

    def __init__(self, index, period_s=0.03):
        import threading
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml as N
            N.nvmlInit()
            self.N = N
            self.h = N.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
            self._sample()                      # first query outside the timed region (NVML warm-up)
            self.samples.clear()
        except Exception:
            self.N = None
            return
        self.period = period_s

    def _sample(self):
        N = self.N
        self.samples.append(float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)))
        r = N.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(N, "nvmlDeviceGetCurrentClocksEventReasons") \
            else N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
                          ("hw_power_brake_slowdown", 0x80)):
            if r & bit:
                self.reasons.add(name)

    def start(self):
        import threading
        if self.N is None:
            return

        def loop():
            while not self._stop.is_set():
                try:
                    self._sample()
                except Exception:
                    break
                self._stop.wait(self.period)
        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)
        sm = self.samples
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm)}


def write_reference_inputs(s, workdir):
    """The same state in the reference's own text formats for oracle/_ref/mh.o."""
    from oracle import oracle as O
    data = O.Data(s["a"], s["d"], s["mu_r"], s["mu_v"], s["n_ssm"], s["cnv_ptr"], s["cnv_idx"], s["cnv_c1"], s["cnv_c2"])
    tree = O.Tree(s["parent"], s["node_of_datum"], s["phi"], s["pi"])
    f = [os.path.join(workdir, n) for n in ("ssm_data.txt", "cnv_data.txt", "c_tree.txt", "c_data_states.txt")]
    O.write_ssm_cnv(data, f[0], f[1])
    O.write_c_tree(tree, f[2], fmt="%.12g")            # py2 str(float) precision, params.py:93
    ssm_of_m, coef = O.data_states(data, tree)
    O.write_c_data_states(ssm_of_m, coef, f[3])
    return f, data, tree


def run_mh_o(files, data, tree, iters, workdir, tag="0", mh_std=MH_STD):
    """One invocation of the reference executable (argv of mh.cpp:22-41).  -> (seconds, acceptance ratio)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "mh.o")
    out = [os.path.join(workdir, "c_params_%s.txt" % tag), os.path.join(workdir, "c_mh_ar_%s.txt" % tag)]
    cmd = [exe, str(iters), str(mh_std), str(data.n_ssm), str(data.n_cnv), str(tree.K), "0"] + files + out + [str(data.T)]
    return cmd, out[1]


def time_reference(s, chains, iters_sample, mh_std=MH_STD):
    """in-loop MH rate of `chains` concurrent mh.o processes: iters/(t(iters) - t(0)); t(0) = the per-call re-parse."""
    exe = os.path.join(ROOT, "oracle", "_ref", "mh.o")
    if not os.path.exists(exe):
        raise SystemExit("oracle/_ref/mh.o missing: run `python __graft_entry__.py` where /root/reference exists")
    with tempfile.TemporaryDirectory(prefix="pwgsdataexchange.") as wd:
        files, data, tree = write_reference_inputs(s, wd)

        def timed(iters):
            t0 = time.perf_counter()
            procs = []
            cores = sorted(os.sched_getaffinity(0))
            for c in range(chains):
                cmd, _ = run_mh_o(files, data, tree, iters, wd, tag=str(c), mh_std=mh_std)
                core = cores[c % len(cores)]                                # one core per chain (N chains = N x 1 core)
                procs.append(subprocess.Popen(cmd, preexec_fn=lambda core=core: os.sched_setaffinity(0, {core})))
            for p in procs:
                if p.wait() != 0:
                    raise SystemExit("reference mh.o failed")
            return time.perf_counter() - t0
        t_parse = min(timed(0), timed(0))
        t_run = timed(iters_sample)
        ar = float(open(os.path.join(wd, "c_mh_ar_0.txt")).read())
    loop = max(t_run - t_parse, 1e-9)
    return dict(rate=chains * iters_sample / loop, t_parse=t_parse, t_run=t_run, accept=ar)


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one MH-kernel launch, from the newest committed `ncu --set full` summary
    under profiles/ (a number measured under the profiler, quoted, never re-measured here); None if there is none."""
    import glob
    import re
    best = None
    def natural(fn):                                                        # r01_v10 after r01_v9
        return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", os.path.basename(fn))]
    for fn in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_mh_kernel_ncu_full.txt")), key=natural):
        rd = wr = None
        for line in open(fn):
            m = re.match(r"dram__bytes_(read|write)\.sum\s+(\S+)\s+(\S+)", line)
            if m:
                v = float(m.group(3)) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(2), float("nan"))
                if m.group(1) == "read":
                    rd = v
                else:
                    wr = v
        if rd is not None and wr is not None:
            best = rd + wr
    return best


def tree_sample_rate(s, device, iterations=12, cpu_mh=None, cpu_seconds=10.0):
    """Secondary figure of BASELINE's metric: full MCMC tree samples/s of one chain through the host glue (evolve.py:148-230:
    assignment sweep + MH + sticks/hypers + LLH trace + pickling) on the same workload, the first iteration discarded.  The
    reference's Python-2 side cannot run here, so there is no CPU number beside it (its sweep is O(N^2))."""
    import pickle
    from oracle import oracle as O                     # only to write the workload in the reference's file formats
    from phylowgs_b200.host import evolve as E
    from phylowgs_b200.host import fastpickle
    from phylowgs_b200.host.backend import GpuBackend
    from phylowgs_b200.host.model import TSSB, alleles, boundbeta
    from phylowgs_b200.host.params import metropolis
    with tempfile.TemporaryDirectory() as wd:
        data = O.Data(s["a"], s["d"], s["mu_r"], s["mu_v"], s["n_ssm"], s["cnv_ptr"], s["cnv_idx"], s["cnv_c1"], s["cnv_c2"])
        ssm, cnv = os.path.join(wd, "ssm_data.txt"), os.path.join(wd, "cnv_data.txt")
        O.write_ssm_cnv(data, ssm, cnv)
        codes, n_ssms, n_cnvs, _ = E.load_data(ssm, cnv)
    E.install_pickle_aliases()
    E.tune_allocator()                                 # as evolve.do_mcmc does
    rng = np.random.RandomState(1)
    backend = GpuBackend(codes, device=device)
    root = alleles(conc=0.1, ntps=len(codes[0].a))
    tssb = TSSB(dp_alpha=25.0, dp_gamma=1.0, alpha_decay=0.25, root_node=root, data=codes, rng=rng, backend=backend)
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
    mh_std, times, parts = MH_STD, [], []
    E.freeze_long_lived()                              # as evolve.do_mcmc does
    for it in range(iterations):
        t = [time.perf_counter()]
        tssb.resample_assignments(); t.append(time.perf_counter())
        tssb.cull_tree()
        if it == 0:
            E.map_datum_to_node(tssb)                  # evolve.do_mcmc: once per run, the sweeps keep datum.node in step afterwards
        acc = float(metropolis(tssb, MH_ITR, mh_std, 0, n_ssms, n_cnvs, "", "", 1, 1, ".", stream=it)); t.append(time.perf_counter())
        if acc < 0.08 and mh_std < 10000:
            mh_std *= 2.0
        if 0.5 < acc < 0.99:
            mh_std /= 2.0
        tssb.resample_sticks(); tssb.resample_stick_orders(); tssb.resample_hypers()
        tssb.complete_data_log_likelihood(); t.append(time.perf_counter())
        fastpickle.dumps(tssb); t.append(time.perf_counter())              # what evolve.do_mcmc calls (protocol-2 pickle of the tree)
        times.append(t[-1] - t[0])
        parts.append(np.diff(t))
    backend.close()
    med = statistics.median(times[1:])
    p = np.median(np.array(parts[1:]), axis=0)
    out = {"value": 1.0 / med, "unit": "tree samples/s", "iterations": iterations - 1, "mh_iterations_per_sample": MH_ITR,
           "seconds": {"assignment_sweep": float(p[0]), "metropolis": float(p[1]), "sticks_hypers_llh": float(p[2]), "pickle": float(p[3])},
           "seconds_per_iteration": [round(x, 5) for x in times],
           "note": "burn-in iterations from the one-node start (many spawned nodes)"}
    if cpu_mh is not None:
        # ---- the CPU beside it ("Python half"): the reference's sweep and LLH trace, restated in plain
        # Python with the reference's cost structure (oracle/py_half.py, held against the reference's own code by
        # tests/test_py_half.py), on the tree the GPU chain has just reached; a bounded sample of the data, scaled to all of them
        from oracle import py_half
        n_data, K_now = tssb.num_data, len(tssb.get_nodes())
        for d in codes:
            d.node = tssb.assignments[d.index]
        t_sw, calls, done, t0 = 0.0, 0, 0, time.perf_counter()
        order = np.random.RandomState(2).permutation(n_data)
        while done < n_data and (done < 4 or time.perf_counter() - t0 < cpu_seconds):
            dt, c = py_half.sweep_sample(tssb, [int(order[done])], rng)
            t_sw += dt
            calls += c
            done += 1
        per_call = t_sw / max(calls, 1)
        sweep_s = t_sw / done * n_data
        llh_s = per_call * n_data                                            # complete_data_log_likelihood: one logprob per datum
        mh_s = cpu_mh["t_parse"] + MH_ITR / cpu_mh["rate"]                   # one mh.o call as params.metropolis makes it
        total = sweep_s + llh_s + mh_s
        out["cpu_baseline"] = {
            "value": 1.0 / total, "unit": "tree samples/s", "cores": 1, "kind": "port",
            "seconds": {"assignment_sweep": sweep_s, "llh_trace": llh_s, "metropolis": mh_s},
            "sample": "oracle/py_half.py (tssb.py:82-150, 404-410 + data.py:26-134 restated; python2 cannot run here): the sweep body "
                      "for %d of the %d data on the %d-node tree the GPU chain reached (%.1f s, %d logprob calls, %.1f ms each: "
                      "every call re-derives the tree metadata, data.py:33-38), scaled to all data; LLH trace = one logprob per "
                      "datum; metropolis = one oracle/_ref/mh.o call of %d iterations at the bench tree's rate incl. its text "
                      "re-parse; stick / hyper-parameter updates and pickling (the same host work on both sides) left out"
                      % (done, n_data, K_now, t_sw, calls, 1e3 * per_call, MH_ITR)}
    return out


def hbm_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except (OSError, ValueError):
        return 6650.0


def reduce_max_over_ranks(x, world, device):
    """Every multi-GPU time is the MAX over ranks (one chain per rank; the job is as slow as its slowest chain)."""
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def chain_stream(rank, phase, k):
    """Philox stream id of step k of a rank's chain in a bench phase (0 warm-up, 1 value, 2 e2e): distinct across ranks, so the
    N chains of a multi-GPU run are independent replicas like multievolve's (multievolve.py:104-105 seeds)."""
    return 1000 * rank + 100 * phase + k


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--mh-batch", type=int, default=0, help="speculative proposals per grid barrier (0 = library default)")
    ap.add_argument("--round-weights", default="", help="tuning: relative cost of a warp round 'plain,cnv1,cnv2,cnv3' for the work "
                                                        "partition (default: the library's)")
    ap.add_argument("--cpu-sample-iters", type=int, default=400, help="MH iterations of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-accept-sweep", action="store_true")
    ap.add_argument("--no-oracle-check", action="store_true")
    ap.add_argument("--no-chain-sweep", action="store_true")
    ap.add_argument("--no-grid", action="store_true")
    ap.add_argument("--no-tree-samples", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="tuning: library option name=value (pwgs_set_option), repeatable")
    ap.add_argument("--mh-std", type=float, default=MH_STD, help="proposal concentration (evolve.py adapts it between 100 and 10000)")
    ap.add_argument("--profile-phases", action="store_true", help="print the MH kernel's per-phase cycle counters (CTA 0) to stderr")
    args = ap.parse_args()
    mh_std = args.mh_std
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, world)

    s, wl_desc = workload(args.workload)
    K = len(s["parent"])
    by_it, _, D, M, T = algorithmic_work(s, K)
    config = {"workload": wl_desc, "n_ssm": int(s["n_ssm"]), "n_cnv": int(D - s["n_ssm"]), "n_cnv_affected_ssm": M, "ntps": T,
              "tree_nodes": K, "mh_iterations_per_metropolis_call": MH_ITR, "mh_std": mh_std, "chains": n_gpus,
              "parallelism": "1 chain per GPU, no collective (timing barrier over gloo)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = []
        acc = None
        for i in range(args.warmup + args.steps):
            r = time_reference(s, n_gpus, args.cpu_sample_iters, mh_std)
            acc = r["accept"]
            if i >= args.warmup:
                steps.append(r)
        rate = statistics.mean(x["rate"] for x in steps)
        ms = 1e3 * statistics.mean(x["t_run"] for x in steps)
        sample = ("%d concurrent oracle/_ref/mh.o (unmodified mh.cpp+util.cpp, g++ -O3, GSL shim), each pinned to its own core, a bounded "
                  "sample of MH_ITR=%d iterations per step on the %s inputs; rate = MH_ITR/(t(MH_ITR)-t(0)), t(0)=%.2fs text re-parse per "
                  "call excluded" % (n_gpus, args.cpu_sample_iters, args.workload, steps[-1]["t_parse"]))
        line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "mh_iterations_per_step": args.cpu_sample_iters, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "accept_ratio": acc,
                "cpu_baseline": {"value": rate, "unit": UNIT, "cores": n_gpus, "kind": "reference", "sample": sample},
                "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    import phylowgs_b200 as P
    if not torch.cuda.is_available() or P.device_count() == 0:
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # The chains share nothing: there is no data-path collective, only the timing barrier and the max-over-ranks of the
        # measured times.  Those are host-side scalars, so the process group is gloo over 127.0.0.1 (no NCCL communicator to set
        # up, and nothing but the JSON line on stdout).
        dist.init_process_group("gloo")
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return reduce_max_over_ranks(x, world, torch.device("cpu"))

    from phylowgs_b200 import synth
    eng = P.Engine(device=local_rank, **synth.data_kwargs(s))
    eng.set_tree(s["parent"], s["node_of_datum"], s["phi"], s["pi"])
    if args.mh_batch:
        eng.set_option("mh_batch", args.mh_batch)
    if args.round_weights:
        for name, v in zip(("mh_w_plain", "mh_w_cnv1", "mh_w_cnv2", "mh_w_cnv3"), args.round_weights.split(",")):
            eng.set_option(name, int(v))
    for kv in args.opt:
        name, v = kv.split("=")
        eng.set_option(name, int(v))
    if args.profile_phases:
        eng.set_option("mh_profile", 1)
    stats = [eng.get_option(n) for n in ("cnv_nu1", "cnv_nu2", "cnv_nu3", "cnv_entry_states")]
    by_it, fp_it, D, M, T = algorithmic_work(s, K, stats)
    ext = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    seed = 20261019

    for w in range(args.warmup):
        with torch.cuda.stream(ext):
            flush.zero_()                                                   # also loads torch's fill kernel before the timed region
        eng.mh_launch(MH_ITR, mh_std, seed, chain_stream(rank, 0, w))
        eng.mh_wait(fetch=False)
    fp64_peak = P.fp64_peak(local_rank)                                 # DFMA thread-instr/s, measured on this GPU now

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # ---- value: device-resident, CUDA events on the library's stream
    launches0 = eng.get_option("launches")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, ratios = [], []
    barrier()
    if sampler:
        sampler.start()
    with torch.cuda.stream(ext):
        e0.record(ext)
        step_wall = []
        for k in range(args.steps):
            tw0 = time.perf_counter()
            flush.zero_()                                               # L2 flush between timed steps (inside the bracket)
            eng.mh_launch(MH_ITR, mh_std, seed, chain_stream(rank, 1, k))
            r = eng.mh_wait(fetch=False)
            kernel_ms.append(r["kernel_ms"])
            ratios.append(r["ratio"])
            step_wall.append(1e3 * (time.perf_counter() - tw0))
        e1.record(ext)
    barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    if rank == 0:
        print("value loop: events %.2f ms total; host wall per step: mean %.2f max %.2f ms; kernel per step: mean %.2f max %.2f ms"
              % (total_ms, statistics.mean(step_wall), max(step_wall), statistics.mean(kernel_ms), max(kernel_ms)), file=sys.stderr)
    launches = eng.get_option("launches") - launches0
    value = n_gpus * args.steps * MH_ITR / (total_ms * 1e-3)
    kms = statistics.mean(kernel_ms)

    # ---- e2e: host arrays in, host arrays out, through the C-ABI calls made once per tree sample
    h2d = s["parent"].nbytes + s["node_of_datum"].nbytes + s["phi"].nbytes + s["pi"].nbytes
    d2h = s["phi"].nbytes + s["pi"].nbytes + 8 + 8
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        eng.set_tree(s["parent"], s["node_of_datum"], s["phi"], s["pi"])
        eng.mh_run(MH_ITR, mh_std, seed, chain_stream(rank, 2, k))
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = n_gpus * args.steps * MH_ITR / e2e_s

    # ---- once, outside the timed regions: the timed workload's result against the CPU oracle (1e-9 relative, north_star)
    oracle_check = None
    if rank == 0 and not args.no_oracle_check:
        from oracle import oracle as O
        od = O.Data(s["a"], s["d"], s["mu_r"], s["mu_v"], s["n_ssm"], s["cnv_ptr"], s["cnv_idx"], s["cnv_c1"], s["cnv_c2"])
        ot = O.Tree(s["parent"], s["node_of_datum"], s["phi"], s["pi"])
        eng.set_tree(s["parent"], s["node_of_datum"], s["phi"], s["pi"])
        res = eng.mh_run(MH_ITR, mh_std, seed, chain_stream(rank, 1, 0))         # the first timed launch again, outputs fetched
        want_entry = O.multi_param_post(od, ot)
        want_final = O.multi_param_post(od, ot, res["phi"], res["pi"])
        got_entry = eng.total_llh(P.SEM_MH)
        rel_entry = abs(got_entry - want_entry) / abs(want_entry)
        rel_final = abs(res["llh"] - want_final) / abs(want_final)
        oracle_check = {"entry_llh_rel_err": rel_entry, "final_llh_rel_err": rel_final, "accepted": int(round(res["ratio"] * MH_ITR)),
                        "tolerance": 1e-9, "what": "pwgs_total_llh at the entry state and the timed launch's final log-likelihood against "
                                                   "oracle multi_param_post (mh.cpp:147-166 restated, pinned to the compiled reference)"}
        if not (rel_entry <= 1e-9 and rel_final <= 1e-9):
            raise SystemExit("bench: the timed workload disagrees with the oracle: %r" % (oracle_check,))

    # ---- the accept path: the same launches at proposal concentrations that make the chain accept (evolve.py:201-206 doubles
    # mh_std up to 10 000 while fewer than 8 % are accepted); kernel time only, CPU beside it at the same mh_std
    accept_sweep = None
    if rank == 0 and world == 1 and not args.no_accept_sweep:
        accept_sweep = []
        for std in [x for x in (100.0, 1.0e4, 1.0e5, 1.0e6) if x != mh_std] + [mh_std]:
            ks, rs = [], []
            for k in range(3 + min(args.steps, 10)):
                with torch.cuda.stream(ext):
                    flush.zero_()
                eng.mh_launch(MH_ITR, std, seed, chain_stream(rank, 4, k))
                r = eng.mh_wait(fetch=False)
                if k >= 3:
                    ks.append(r["kernel_ms"])
                    rs.append(r["ratio"])
            its = MH_ITR / (statistics.mean(ks) * 1e-3)
            row = {"mh_std": std, "value": its, "unit": UNIT, "kernel_ms_per_launch": statistics.mean(ks), "accept_ratio": statistics.mean(rs),
                   "roofline_frac": fp_it * its / fp64_peak}
            if not args.no_cpu_baseline:
                rr = time_reference(s, 1, max(100, args.cpu_sample_iters // 2), std)
                row["cpu_baseline"] = {"value": rr["rate"], "unit": UNIT, "cores": 1, "kind": "reference", "accept_ratio": rr["accept"]}
            accept_sweep.append(row)
        accept_sweep.sort(key=lambda r: r["mh_std"])

    # ---- 1-8 chains of this workload on ONE GPU (BASELINE.json configs[4]): one cooperative launch, 148 / C SMs per chain
    chains_per_gpu = None
    if rank == 0 and world == 1 and not args.no_chain_sweep:
        chains_per_gpu = []
        extra = []
        try:
            for C_ in (1, 2, 4, 8):
                while len(extra) < C_ - 1:
                    e2 = P.Engine(device=local_rank, **synth.data_kwargs(s))
                    e2.set_tree(s["parent"], s["node_of_datum"], s["phi"], s["pi"])
                    for kv in args.opt:
                        e2.set_option(kv.split("=")[0], int(kv.split("=")[1]))
                    extra.append(e2)
                engs = [eng] + extra[:C_ - 1]
                ks = []
                for k in range(3 + min(args.steps, 10)):
                    with torch.cuda.stream(ext):
                        flush.zero_()
                    P.Engine.mh_launch_multi(engs, MH_ITR, mh_std, seed, [chain_stream(c_, 5, k) for c_ in range(C_)])
                    rr = [e_.mh_wait(fetch=False) for e_ in engs]
                    if k >= 3:
                        ks.append(max(r_["kernel_ms"] for r_ in rr))
                chains_per_gpu.append({"chains": C_, "sms_per_chain": eng.get_option("n_sms") // C_, "value": C_ * MH_ITR / (statistics.mean(ks) * 1e-3),
                                       "unit": UNIT, "kernel_ms_per_launch": statistics.mean(ks),
                                       # the same instruction model as `roofline`, all chains of the launch together
                                       "roofline_frac": (fp_it * C_ * MH_ITR / (statistics.mean(ks) * 1e-3) / fp64_peak) if fp_it else None})
        finally:
            for e2 in extra:
                e2.close()

    # ---- the grid kernel K2 (SURVEY.md 8d "Grid kernel K2"): every datum in every node of a 128-node tree over the same data, the
    # size an assignment sweep asks for at this workload (bench.py's tree samples run at 100-200 nodes); device time from the
    # library's own events around the kernel, end-to-end time with the grid landing in page-locked host memory
    grid_kernel = None
    if rank == 0 and world == 1 and not args.no_grid:
        Kg, rng_g = 128, np.random.default_rng(1)
        parent_g = np.array([-1] + [int(rng_g.integers(0, k)) for k in range(1, Kg)], dtype=np.int32)
        pi_g = rng_g.dirichlet(np.ones(Kg), size=T).T.copy()
        phi_g = pi_g.copy()
        for k in range(Kg - 1, 0, -1):
            phi_g[parent_g[k]] += phi_g[k]
        nod_g = rng_g.integers(1, Kg, size=D).astype(np.int32)
        eng.set_tree(parent_g, nod_g, phi_g, pi_g)
        buf = eng.host_buffer((Kg, D))
        ks, es = [], []
        for k in range(8):
            t0 = time.perf_counter()
            eng.llh_range(0, D, None, P.SEM_PY, col_major=True, out=buf)
            es.append(time.perf_counter() - t0)
            ks.append(eng.get_option("grid_kernel_ns") * 1e-9)
        k_s, e_s = statistics.median(ks[2:]), statistics.median(es[2:])
        pairs = D * Kg
        n_plain_pairs, n_cnv_pairs = (D - M) * Kg, M * Kg
        # FP64-pipe instructions of the kernel as compiled (libm log / exp: profiles/r02_pair_llh_kernel_sass_histogram.txt):
        # a plain pair = 2 logs (29 each) + 8; a CNV-affected pair (data.py convention) ~ 2-4 states x (2 logs + division 13 + 8)
        # + logsumexp (exp 16 each + log 29) ~ 400 per sample
        fp_grid = T * (n_plain_pairs * 66.0 + n_cnv_pairs * 400.0)
        by_grid = D * (20 + 8 * T) + 8.0 * pairs
        eng.set_tree(s["parent"], s["node_of_datum"], s["phi"], s["pi"])
        grid_kernel = {"pairs": pairs, "tree_nodes": Kg, "kernel_ms": 1e3 * k_s, "e2e_ms": 1e3 * e_s, "pairs_per_s": pairs / k_s,
                       "d2h_bytes": 8 * pairs,
                       "roofline": {"bound": "fp64", "achieved": fp_grid / k_s / 1e9, "peak": fp64_peak / 1e9, "unit": "G FP64-instr/s",
                                    "frac": fp_grid / k_s / fp64_peak,
                                    "hbm": {"achieved": by_grid / k_s / 1e9, "peak": hbm_peak_gbs(), "unit": "GB/s", "frac": by_grid / k_s / 1e9 / hbm_peak_gbs()},
                                    "pcie_d2h_GBps_e2e": 8 * pairs / e_s / 1e9}}

    if args.profile_phases and rank == 0:
        names = ["likelihood", "publish", "speculative_draw", "barrier_wait", "reduce", "decide", "accept_path", "batches"]
        pc = [eng.get_option("mh_prof_%d" % i) for i in range(8)]
        tot = max(sum(pc[:7]), 1)
        print("phase cycles (CTA 0, last launch): " + ", ".join("%s=%d (%.1f%%)" % (n, v, 100.0 * v / tot) for n, v in zip(names[:7], pc[:7]))
              + ", batches=%d, cycles/batch=%.0f" % (pc[7], tot / max(pc[7], 1)), file=sys.stderr)
        lo, hi, sm = (eng.get_option("mh_lik_" + k) for k in ("min", "max", "sum"))
        ncta = eng.get_option("mh_ctas") or eng.get_option("n_sms")
        print("likelihood cycles over the chain's CTAs: min %d  mean %.0f  max %d  (max/mean %.3f)" % (lo, sm / ncta, hi, hi * ncta / max(sm, 1)),
              file=sys.stderr)
    # ---- opt-in fast path, reported separately (not the headline): plain data through exact integer totals (K0d)
    collapsed = None
    if eng.get_option("n_classes") > 0:
        eng.set_option("mh_collapse", 1)
        cms = []
        for k in range(3 + min(args.steps, 10)):
            eng.mh_launch(MH_ITR, mh_std, seed, chain_stream(rank, 3, k))
            r = eng.mh_wait(fetch=False)
            if k >= 3:
                cms.append(r["kernel_ms"])
        eng.set_option("mh_collapse", 0)
        collapsed = {"value": n_gpus * MH_ITR / (statistics.mean(cms) * 1e-3), "unit": UNIT, "kernel_ms_per_launch": statistics.mean(cms),
                     "note": "mh_collapse=1: plain data enter as integer totals per (node, error-rate class, sample); exact, opt-in, "
                             "kernel time only, not comparable with the roofline above (different algorithmic work)"}
    if rank != 0:
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return 0
    tree_samples = None
    if world == 1 and not args.no_tree_samples:
        cpu_mh = None if args.no_cpu_baseline else time_reference(s, 1, args.cpu_sample_iters, mh_std)
        tree_samples = tree_sample_rate(s, local_rank, cpu_mh=cpu_mh)
    else:
        cpu_mh = None

    # ---- roofline of the dominant kernel (mh_kernel), per launch
    its_kernel = MH_ITR / (kms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = {"bound": "fp64", "achieved": fp_it * its_kernel / 1e9, "peak": fp64_peak / 1e9, "unit": "G FP64-instr/s",
                "frac": fp_it * its_kernel / fp64_peak, "traffic": ncu_traffic_bytes() if args.workload == "config3" else None,
                "peak_source": "pwgs_fp64_peak: DFMA chains measured live on this GPU (MEASURED_PEAKS.json has no FP64 figure)",
                "algorithmic_fp64_instr_per_iteration": fp_it, "kernel_ms_per_launch": kms,
                "frac_on_round1_instruction_model": algorithmic_work_r1(s, stats) * its_kernel / fp64_peak,
                "round1_model_fp64_instr_per_iteration": algorithmic_work_r1(s, stats),
                # SURVEY.md 8d's own per-unit figure (the reference's dense 4-state formulation with libm: 66 per plain term,
                # 8K + 400 per CNV-affected SSM): > 1 means the kernel gets the iteration done with less arithmetic than that
                # formulation needs at the pipe's peak -- an algorithmic gain (sparse merged states, table-driven log), not utilisation
                "survey_8d_model_fp64_instr_per_iteration": float(T * (66 * (D - M) + (8 * K + 400) * M)),
                "frac_on_survey_8d_model": float(T * (66 * (D - M) + (8 * K + 400) * M)) * its_kernel / fp64_peak,
                "note": "frac = FP64-pipe instructions this cubin executes per completed iteration (profiles/r02_mh_kernel_sass_histogram.txt: "
                        "9 per logarithm) x it/s / measured DFMA peak = the pipe utilisation ncu reports; the round-1 model (16 per "
                        "logarithm) is the same work counted the way BENCH_r01 counted it",
                "cnv_ssm_by_merged_states": stats[:3], "cnv_entry_states": stats[3],
                "hbm": {"achieved": by_it * its_kernel / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": by_it * its_kernel / 1e9 / hbm_peak,
                        "algorithmic_bytes_per_iteration": by_it,
                        "note": "what streaming every record from HBM once per iteration would need; the kernel reads them once per "
                                "LAUNCH and keeps them in shared memory (see traffic), which is why HBM is not the bound",
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_mh or time_reference(s, 1, args.cpu_sample_iters, mh_std)
        cpu_baseline = {"value": r["rate"], "unit": UNIT, "cores": 1, "kind": "reference",
                        "sample": "oracle/_ref/mh.o (unmodified mh.cpp+util.cpp, g++ -O3, GSL shim) MH_ITR=%d on the same %s inputs: %.2fs total, "
                                  "%.2fs of it the per-call text re-parse (excluded from the rate); accept=%.4f; host has %d cores"
                                  % (args.cpu_sample_iters, args.workload, r["t_run"], r["t_parse"], r["accept"], os.cpu_count())}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "mh_iterations_per_step": MH_ITR, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "timing": {"l2": "256 MiB L2 flush enqueued between timed steps; working set is on-chip by design", "mh_batch": eng.get_option("mh_batch")},
            "accept_ratio": statistics.mean(ratios), "clocks": clocks, "oracle_check": oracle_check, "accept_sweep": accept_sweep, "chains_per_gpu": chains_per_gpu, "grid_kernel": grid_kernel,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "collapsed_plain_data": collapsed,
            "tree_samples": tree_samples}
    print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
