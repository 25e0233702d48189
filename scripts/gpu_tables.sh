# Numbers for DESIGN.md's tables (one B200): acceptance sweep, other sizes, K2, config 4.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for std in 100 10000 100000 1000000; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mh-std $std 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('std $std it/s %.0f'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'acc %.5f'%d['accept_ratio'], 'collapsed', d.get('collapsed_plain_data',{}).get('kernel_ms_per_launch'))"; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload config4 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('config4 it/s %.0f'%d['value'], 'ms %.3f'%d['roofline']['kernel_ms_per_launch'], 'frac %.3f'%d['roofline']['frac'], 'collapsed', d.get('collapsed_plain_data',{}).get('kernel_ms_per_launch'))"
timeout 300 python scripts/small_case.py 2>&1 | tail -6
timeout 600 python scripts/big_case.py 2>&1 | tail -8
timeout 300 python scripts/k2_time.py 2>&1 | tail -6
