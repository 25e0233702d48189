# Quick GPU iteration: parity tests + short bench at several batch caps.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for b in 16 4 1; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mh-batch $b --profile-phases 2>gpurun_out/err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('batch', d['config']['mh_batch'], 'it/s %.0f'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.0f'%d['e2e']['value'], 'acc', d['accept_ratio'], d['clocks'])"; tail -3 gpurun_out/err.txt; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mh-std 100000 --profile-phases 2>gpurun_out/err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('std1e5 batch', d['config']['mh_batch'], 'it/s %.0f'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.0f'%d['e2e']['value'], 'acc', d['accept_ratio'])"; tail -3 gpurun_out/err.txt
