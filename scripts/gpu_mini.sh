# Shortest GPU check of a K1 change: trajectory parity tests + the default bench with phase counters.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-phases 2>gpurun_out/err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('it/s %.0f'%d['value'], 'kernel ms %.4f'%d['roofline']['kernel_ms_per_launch'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.0f'%d['e2e']['value'], 'acc', d['accept_ratio'])"
tail -2 gpurun_out/err.txt
