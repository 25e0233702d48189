# MH it/s against the number of CTAs a chain runs on, for small inputs at high acceptance (where chain barriers dominate).
for c in ${CASES:-config1:1000000 config1:100 config5:1000:100000 config5:1000:100 config2:1000000 config2:100}; do
  w=${c%:*}; std=${c##*:}
  for n in 1 2 4 8 16 37 148; do
    timeout 300 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline --no-accept-sweep --no-oracle-check --no-chain-sweep --no-grid --no-tree-samples --mh-std $std --opt mh_ctas=$n 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$w std=$std ctas=$n: kernel_ms=%.4f value=%.0f accept=%.4f' % (d['roofline']['kernel_ms_per_launch'], d['value'], d['accept_ratio']))"
  done
done
