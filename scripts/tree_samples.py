"""Times the phases of full MCMC tree samples (evolve.py:148-230) of the host glue on the GPU backend:
   python scripts/tree_samples.py <ssm> <cnv|-> [iterations] [mh_iters]      ('config3' as <ssm> = synthetic WGS tumour)"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import chain_util
from phylowgs_b200.host import evolve as E
from phylowgs_b200.host.backend import GpuBackend
from phylowgs_b200.host.params import metropolis

ssm, cnv = sys.argv[1], sys.argv[2]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mh_iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5000
tmp = tempfile.mkdtemp()
if ssm == "config3":
    from oracle import oracle as O          # only to write the synthetic tumour in the reference's file formats
    from phylowgs_b200 import synth
    s = synth.config(3)
    data = O.Data(s["a"], s["d"], s["mu_r"], s["mu_v"], s["n_ssm"], s["cnv_ptr"], s["cnv_idx"], s["cnv_c1"], s["cnv_c2"])
    ssm, cnv = os.path.join(tmp, "ssm.txt"), os.path.join(tmp, "cnv.txt")
    O.write_ssm_cnv(data, ssm, cnv)
if cnv == "-":
    cnv = os.path.join(tmp, "empty.txt"); open(cnv, "w").close()
t0 = time.time()
tssb, n_ssms, n_cnvs = chain_util.new_chain(ssm, cnv, lambda c: GpuBackend(c, 0), seed=1)
print("load + context: %.2fs  (%d SSMs, %d CNVs, T=%d)" % (time.time() - t0, n_ssms, n_cnvs, len(tssb.data[0].a)))
mh_std = 100.0
import pickle
from phylowgs_b200.host import fastpickle
E.tune_allocator(); E.freeze_long_lived()
E.install_pickle_aliases()
for it in range(iters):
    t = [time.time()]
    tssb.resample_assignments(); t.append(time.time())
    tssb.cull_tree()
    _, nodes = tssb.get_mixture()
    for i, nd in enumerate(nodes): nd.id = i
    E.set_node_height(tssb); E.set_path_from_root_to_node(tssb); E.map_datum_to_node(tssb); t.append(time.time())
    acc = float(metropolis(tssb, mh_iters, mh_std, 0, n_ssms, n_cnvs, "", "", 1, 1, ".", stream=it)); t.append(time.time())
    if acc < 0.08 and mh_std < 10000: mh_std *= 2
    if 0.5 < acc < 0.99: mh_std /= 2
    tssb.resample_sticks(); tssb.resample_stick_orders(); tssb.resample_hypers(); t.append(time.time())
    llh = tssb.complete_data_log_likelihood(); t.append(time.time())
    blob = fastpickle.dumps(tssb); t.append(time.time())
    d = np.diff(t)
    print("it %2d: total %.3fs | assign %.3f (grid refreshes %d) meta %.3f mh %.4f (acc %.3f std %g) sticks/hypers %.3f llh %.4f pickle %.3f (%d KB) | nodes %d llh %.1f"
          % (it, t[-1] - t[0], d[0], tssb.last_grid_refreshes, d[1], d[2], acc, mh_std, d[3], d[4], d[5], len(blob) // 1024, len(nodes), llh), flush=True)
