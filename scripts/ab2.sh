#!/bin/bash
# A/B of kernel variants over workloads and proposal widths: scripts/ab2.sh lib1.so lib2.so ...   (AB_CASES="config2:1000000 config3:100 ...")
for lib in "$@"; do
  tag=$(basename "$lib" .so)
  echo "=== $tag"
  PWGS_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q 2>&1 | tail -1
  for c in ${AB_CASES:-config3:100 config3:100000 config2:100000 config2:1000000 config1:1000000}; do
    w=${c%%:*}; std=${c##*:}
    PWGS_LIB=$PWD/$lib timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-accept-sweep --no-oracle-check --no-chain-sweep --no-grid --no-tree-samples --mh-std $std 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('  $w mh_std=%s kernel_ms=%.4f value=%.0f accept=%.5f' % (d['config']['mh_std'], d['roofline']['kernel_ms_per_launch'], d['value'], d['accept_ratio']))"
  done
done
