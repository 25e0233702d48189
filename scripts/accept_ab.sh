# Accept-path A/B: kernel time per 5000 iterations over proposal widths on configs 1 / 2 / 3 (bench lines' accept_sweep).
for w in config3 config2 config1; do
  timeout 600 python bench.py --workload $w --no-tree-samples --no-chain-sweep --no-grid --no-cpu-baseline --steps 10 > gpurun_out/ab_$w.json 2>> gpurun_out/ab.err
  python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/ab_$w.json") if l.startswith("{")][-1])
print("$w", "value %.0f kernel %.4f ms |" % (j["value"], j["roofline"]["kernel_ms_per_launch"]),
      " ".join("std=%g acc=%.4f %.0f it/s (%.3f ms)" % (r["mh_std"], r["accept_ratio"], r["value"], 5000e3 / r["value"]) for r in j["accept_sweep"]))
PY
done
