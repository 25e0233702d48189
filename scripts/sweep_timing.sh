# Where a tree sample's assignment sweep spends its time (config 3, burn-in iterations from the one-node start).
PWGS_SWEEP_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 3 2>&1 >/dev/null | grep "sweep timing"
