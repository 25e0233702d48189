#!/bin/bash
# A/B timing of kernel variants on the GPU box: for each library file given (built with phylowgs_b200/build.py -D.. -o..),
# the MH trajectory parity tests, then the bench's kernel time and phase counters.  Usage: scripts/ab.sh lib1.so lib2.so ...
mkdir -p gpurun_out
for lib in "$@"; do
  tag=$(basename "$lib" .so)
  echo "=== $tag"
  PWGS_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
  for std in ${AB_STDS:-100}; do
    PWGS_LIB=$PWD/$lib timeout 300 python bench.py --steps ${AB_STEPS:-10} --warmup 3 --no-cpu-baseline --no-accept-sweep --no-oracle-check --no-chain-sweep --no-grid --profile-phases --mh-std $std ${AB_ARGS} \
      2> gpurun_out/ab_${tag}_$std.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('  mh_std=%s kernel_ms=%.4f value=%.0f e2e=%.0f frac=%.4f accept=%.5f' % (d['config']['mh_std'], d['roofline']['kernel_ms_per_launch'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['accept_ratio']))"
    grep -E "phase cycles|likelihood cycles" gpurun_out/ab_${tag}_$std.err | sed 's/^/  /'
  done
done
