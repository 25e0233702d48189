# Steady-state tree samples/s of n chains sharing ONE GPU, from the chains' own mcmc_samples.txt (per-sample seconds; start-up
# excluded): one evolve process per chain against --in-process.     bash scripts/chain_rate.sh [ssm] [n_list]
SSM=${1:-tests/golden/emptysim.6.300.2000.ssm.txt}
touch /tmp/empty_cnv.txt
for mode in "" "--in-process"; do
for n in ${2:-1 4 8}; do
  rm -rf /tmp/cr_$n
  python -m phylowgs_b200.host.multievolve -n $n --gpus 0 $mode -O /tmp/cr_$n --ssms $SSM --cnvs /tmp/empty_cnv.txt -B 100 -s 200 > /tmp/cr_$n.log 2>&1 || tail -5 /tmp/cr_$n.log
  python - <<PY
import glob, numpy as np
rates = []
for fn in sorted(glob.glob("/tmp/cr_$n/chain_*/mcmc_samples.txt")):
    dt = np.array([float(l.split("\t")[2]) for l in list(open(fn))[1:] if l.strip()])[50:]      # skip the first samples (growing tree, warm-up)
    rates.append(len(dt) / dt.sum())
print("chains on one GPU %-12s: %d  -> %.1f tree samples/s aggregate in steady state (%.1f per chain; 5000 MH iterations per sample)" % ("$mode" or "(processes)", $n, sum(rates), np.mean(rates) if rates else 0))
PY
done
done
