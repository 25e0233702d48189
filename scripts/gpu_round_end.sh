# End-of-round evidence: gpu_full.sh (tests, smoke, both bench arms, launch list, full capture of K1) + the bench lines of the other
# named configurations + one full capture of the grid kernel K2 at the size a sweep asks for (50 300 data x 120 nodes).
TAG=${1:-r02_v5}
bash scripts/gpu_full.sh $TAG > gpurun_out/${TAG}_full.log 2>&1
tail -c 600 gpurun_out/${TAG}_full.log
for w in config1 config2 config4 config5:1000 config5:100000 config5:500000; do
  timeout 900 python bench.py --workload $w --no-tree-samples > gpurun_out/${TAG}_bench_${w/:/_}.json 2>> gpurun_out/bench.err
done
timeout 300 python scripts/k2_time.py > gpurun_out/k2_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_llh_sorted -s 8 -c 1 -f -o gpurun_out/${TAG}_prof_k2 python scripts/k2_time.py > gpurun_out/ncu_k2.log 2>&1
tail -3 gpurun_out/k2_plain.log; tail -2 gpurun_out/ncu_k2.log; ls -la gpurun_out/
