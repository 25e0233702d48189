# One ncu --set full capture of the MH kernel (after the same command exited 0 without ncu).
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-accept-sweep --no-oracle-check --no-chain-sweep --no-grid > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:mh_kernel -s 1 -c 1 -f -o gpurun_out/prof_mh python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-accept-sweep --no-oracle-check --no-chain-sweep --no-grid > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; cat gpurun_out/plain.log | tail -1 | cut -c1-300
