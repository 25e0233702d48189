# Aggregate tree samples/s of n independent chains sharing ONE GPU: one evolve process per chain (multievolve's default,
# round-robin placement) against --in-process (chains = threads of one process, one CUDA context).
SSM=tests/golden/emptysim.6.300.2000.ssm.txt
touch /tmp/empty_cnv.txt
for mode in "" "--in-process"; do
for n in 1 4 8; do
  rm -rf /tmp/cpg_$n
  t0=$(date +%s.%N)
  python -m phylowgs_b200.host.multievolve -n $n --gpus 0 $mode -O /tmp/cpg_$n --ssms $SSM --cnvs /tmp/empty_cnv.txt -B 60 -s 60 > /tmp/cpg_$n.log 2>&1 || tail -5 /tmp/cpg_$n.log
  t1=$(date +%s.%N)
  python - <<PY
n=$n; dt=$t1-$t0
print("chains on one GPU %-12s: %d  wall %.1f s  -> %.1f tree samples/s aggregate (120 samples x 5000 MH iterations per chain, incl. start-up, pickling, zip)" % ("$mode" or "(processes)", n, dt, n*120/dt))
PY
done
done
