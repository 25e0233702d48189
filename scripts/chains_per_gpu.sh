# Aggregate tree samples/s of n independent evolve processes sharing ONE GPU (multievolve round-robin placement).
SSM=tests/golden/emptysim.6.300.2000.ssm.txt
touch /tmp/empty_cnv.txt
for n in 1 4 8; do
  rm -rf /tmp/cpg_$n
  t0=$(date +%s.%N)
  python -m phylowgs_b200.host.multievolve -n $n --gpus 0 -O /tmp/cpg_$n --ssms $SSM --cnvs /tmp/empty_cnv.txt -B 60 -s 60 > /tmp/cpg_$n.log 2>&1
  t1=$(date +%s.%N)
  python - <<PY
n=$n; dt=$t1-$t0
print("chains on one GPU: %d  wall %.1f s  -> %.1f tree samples/s aggregate (120 samples x 5000 MH iterations per chain, incl. process start-up, pickling, zip)" % (n, dt, n*120/dt))
PY
done
