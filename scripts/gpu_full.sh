# Full single-GPU pass: parity tests, smoke, bench (both arms), ncu launch list + one full capture of the MH kernel.
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
echo "host: $(nproc) cores, $(lscpu | grep 'Model name' | sed 's/.*: *//')"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/${TAG}_bench_reference.json
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/bench.err; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-accept-sweep --no-oracle-check --no-chain-sweep --no-grid > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-accept-sweep --no-oracle-check --no-chain-sweep --no-grid > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mh_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_mh python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-accept-sweep --no-oracle-check --no-chain-sweep --no-grid > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
