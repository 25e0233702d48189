# Final check of a round: GPU tests, smoke, bench (both arms) -> gpurun_out/<tag>_bench*.json  (no profiler)
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('value %.0f e2e %.0f frac %.3f kernel_ms %.4f cpu %.1f tree_samples %s clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], d['cpu_baseline']['value'], d['tree_samples'], d['clocks']))"
