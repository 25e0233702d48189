"""End-to-end inference on a phylowgs_sims file with the GPU chain: posterior-mean prevalence per simulated block vs 2*VAF."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import chain_util
from phylowgs_b200.host.backend import GpuBackend
from phylowgs_b200.host.params import metropolis
ssm = sys.argv[1]; burn = int(sys.argv[2]); keep = int(sys.argv[3]); block = int(sys.argv[4])
tmp = tempfile.mkdtemp(); cnv = os.path.join(tmp, "e.txt"); open(cnv, "w").close()
tssb, n_ssms, n_cnvs = chain_util.new_chain(ssm, cnv, lambda c: GpuBackend(c, 0), seed=1)
rows = [l.split("\t") for l in open(ssm).read().strip().split("\n")[1:]]
a = np.array([int(r[2]) for r in rows], float); d = np.array([int(r[3]) for r in rows], float)
nb = n_ssms // block
truth = np.array([2 * (1 - a[i*block:(i+1)*block] / d[i*block:(i+1)*block]).mean() for i in range(nb)])
phi_sum = np.zeros(n_ssms); major = []; mh_std = 100.0; t0 = time.time(); llhs = []
for it in range(burn + keep):
    tssb.resample_assignments(); tssb.cull_tree()
    acc = float(metropolis(tssb, 5000, mh_std, 0, n_ssms, n_cnvs, "", "", 1, 1, ".", stream=it))
    if acc < 0.08 and mh_std < 10000: mh_std *= 2
    if 0.5 < acc < 0.99: mh_std /= 2
    tssb.resample_sticks(); tssb.resample_stick_orders(); tssb.resample_hypers()
    llhs.append(tssb.complete_data_log_likelihood())
    if it >= burn:
        phi_sum += np.array([float(tssb.assignments[n].params[0]) for n in range(n_ssms)])
        major.append(sum(1 for nd in tssb.get_nodes() if nd.num_local_data() >= 0.02 * n_ssms))
dt = time.time() - t0
post = phi_sum / keep
print("%s: %d SSMs, %d+%d tree samples x 5000 MH iterations in %.1f s (%.1f tree samples/s)" % (os.path.basename(ssm), n_ssms, burn, keep, dt, (burn + keep) / dt))
print("block truth 2*VAF :", np.round(truth, 3))
print("posterior mean phi:", np.round([post[i*block:(i+1)*block].mean() for i in range(nb)], 3))
print("per-SSM RMSE vs block truth: %.4f ; populations holding >= 2%% of the SSMs (mode): %d ; LLH first/last: %.1f / %.1f"
      % (np.sqrt(np.mean((post - np.repeat(truth, block)) ** 2)), np.bincount(major).argmax(), llhs[0], llhs[-1]))
