bash scripts/gpu_mini.sh
python scripts/partition_fit.py config4 100,160,400,640 -v
for std in 10000 100000; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mh-std $std 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('std $std it/s %.0f'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'acc', d['accept_ratio'])"; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload config4 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('config4 it/s %.0f'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'acc', d['accept_ratio'], d['config'])"
