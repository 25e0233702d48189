"""Where one params.metropolis() call goes on the trees of a running chain (config 3, ~100 nodes during burn-in): host flattening,
pwgs_set_tree, the kernel (library's own events) and its phases (mh_profile).   python scripts/mh_tree_profile.py [tree samples]"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import chain_util
from phylowgs_b200.host import evolve as E
from phylowgs_b200.host.backend import GpuBackend, flatten_tree
from oracle import oracle as O          # only to write the synthetic tumour in the reference's file formats
from phylowgs_b200 import synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
tmp = tempfile.mkdtemp()
s = synth.config(3)
data = O.Data(s["a"], s["d"], s["mu_r"], s["mu_v"], s["n_ssm"], s["cnv_ptr"], s["cnv_idx"], s["cnv_c1"], s["cnv_c2"])
ssm, cnv = os.path.join(tmp, "ssm.txt"), os.path.join(tmp, "cnv.txt")
O.write_ssm_cnv(data, ssm, cnv)
tssb, n_ssms, n_cnvs = chain_util.new_chain(ssm, cnv, lambda c: GpuBackend(c, 0), seed=1)
eng = tssb.backend.engine
eng.set_option("mh_profile", 1)
names = ["likelihood", "publish", "speculative_draw", "barrier_wait", "reduce", "decide", "accept_path"]
mh_std = 100.0
T = time.perf_counter
for it in range(iters):
    tssb.resample_assignments()
    tssb.cull_tree()
    t0 = T()
    nodes, parent, nod, phi, pi = flatten_tree(tssb); t1 = T()
    eng.set_tree(parent, nod, phi, pi); t2 = T()
    eng.mh_launch(5000, mh_std, 1, it); t3 = T()
    r = eng.mh_wait(); t4 = T()
    for k, nd in enumerate(nodes):
        nd.params = np.array(r["phi"][k]); nd.pi = np.array(r["pi"][k])
    t5 = T()
    acc = r["ratio"]
    pc = [eng.get_option("mh_prof_%d" % i) for i in range(8)]
    tot = max(sum(pc[:7]), 1)
    print("it %d K=%d std=%g acc=%.4f | flatten %.2f set_tree %.2f launch(host: work list, plan) %.2f wait %.2f [kernel %.3f] write-back %.2f ms | %s batches=%d"
          % (it, len(nodes), mh_std, acc, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), r["kernel_ms"], 1e3 * (t5 - t4),
             " ".join("%s=%.0f%%" % (n, 100.0 * v / tot) for n, v in zip(names, pc[:7])), pc[7]), flush=True)
    if acc < 0.08 and mh_std < 10000: mh_std *= 2
    if 0.5 < acc < 0.99: mh_std /= 2
    tssb.resample_sticks(); tssb.resample_stick_orders(); tssb.resample_hypers()
