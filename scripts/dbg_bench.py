import sys, time, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import phylowgs_b200 as P
from phylowgs_b200 import synth
import bench
s = synth.config(3)
eng = P.Engine(device=0, **synth.data_kwargs(s))
eng.set_tree(s["parent"], s["node_of_datum"], s["phi"], s["pi"])
dev = torch.device("cuda", 0)
ext = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for w in range(3):
    eng.mh_launch(5000, 100.0, 1, w); eng.mh_wait(fetch=False)
def run(n, use_flush, use_sampler, period=0.15):
    samp = bench.ClockSampler(0, period) if use_sampler else None
    if samp: samp.start()
    tf = tl = tw = 0.0; kms = 0.0
    torch.cuda.synchronize()
    t00 = time.perf_counter()
    with torch.cuda.stream(ext):
        for k in range(n):
            t0 = time.perf_counter()
            if use_flush: flush.zero_()
            t1 = time.perf_counter()
            eng.mh_launch(5000, 100.0, 1, 100 + k)
            t2 = time.perf_counter()
            r = eng.mh_wait(fetch=False)
            t3 = time.perf_counter()
            tf += t1 - t0; tl += t2 - t1; tw += t3 - t2; kms += r["kernel_ms"]
    tot = time.perf_counter() - t00
    c = samp.stop() if samp else None
    print("n=%d flush=%d sampler=%d: total %.1f ms/step  flush %.2f launch %.2f wait %.2f kernel %.2f" % (n, use_flush, use_sampler, 1e3*tot/n, 1e3*tf/n, 1e3*tl/n, 1e3*tw/n, kms/n), c)
run(20, 0, 0); run(20, 1, 0); run(20, 1, 1); run(50, 1, 1); run(50, 0, 1); run(50, 1, 0); run(50,1,1,0.5)
