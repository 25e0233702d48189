# Calibrate the work partition's round weights on config 3 (plain fixed at 100).
mkdir -p gpurun_out
for w1 in 110 130 150 175; do for w2 in 280 330 380 430 500; do
  timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --round-weights 100,$w1,$w2,$((w2*3/2)) --profile-phases 2>gpurun_out/err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('w1 $w1 w2 $w2  kernel ms %.4f'%d['roofline']['kernel_ms_per_launch'])"
  grep -o "barrier_wait=[0-9]* ([0-9.]*%)" gpurun_out/err.txt
done; done
