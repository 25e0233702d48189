#!/bin/bash
# End-to-end statistical comparison, GPU side: 4 seeds per set with this repo's evolve driver on cuda:0, summarised next to the
# frozen reference summaries (tests/golden/stat_ref_<set>.json, made here by scripts/stat_parity.py ref + summarise).
mkdir -p gpurun_out
for set in ${STAT_SETS:-config2 config1 sim150}; do
  rm -rf /tmp/gs_$set; mkdir -p /tmp/gs_$set
  for s in ${STAT_SEEDS:-1 2 3 4}; do
    ( timeout 1200 python scripts/stat_parity.py gpu --set $set --seed $s -O /tmp/gs_$set/s$s > /tmp/gs_$set/s$s.out 2>&1 || echo "chain $set $s FAILED" ) &
  done
  wait
  python scripts/stat_parity.py summarise $(for s in ${STAT_SEEDS:-1 2 3 4}; do echo /tmp/gs_$set/s$s; done) -o gpurun_out/stat_gpu_$set.json
  if [ -f tests/golden/stat_ref_$set.json ]; then python scripts/stat_parity.py compare tests/golden/stat_ref_$set.json gpurun_out/stat_gpu_$set.json | tee gpurun_out/stat_compare_$set.md; fi
done
