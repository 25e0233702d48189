# cProfile of one evolve run on the 10k-SSM sim: where the host time of a tree sample goes.
touch /tmp/empty_cnv.txt; rm -rf /tmp/prof_run
python - <<'PY'
import cProfile, pstats, sys, threading, io
sys.argv = ["evolve", "-O", "/tmp/prof_run", "-B", "100", "-s", "100", "-r", "3", "tests/golden/emptysim.6.300.2000.ssm.txt", "/tmp/empty_cnv.txt"]
from phylowgs_b200.host import evolve as E
pr = cProfile.Profile()
safe, ok = threading.Event(), threading.Event()
import time
t0 = time.time()
pr.enable()
E.run(safe, ok, {"tmp_dir": None}, sys.argv[1:])
pr.disable()
print("wall %.2f s for 200 samples, succeeded=%s" % (time.time() - t0, ok.is_set()))
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:6000])
PY
