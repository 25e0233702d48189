"""Freezes the reference's write_results outputs (write_results.py:44-55: ResultGenerator -> ResultMunger -> IndexCalculator /
calc_tree_densities, executed under Python 3 by oracle/ref_py3.py) for two hand-shaped archives:

    python tests/golden/make_results_golden.py      (build container only: needs /root/reference)

  results_bundled.zip   12 trees over the bundled example (11 SSMs + 1 CNV x 5 samples)
  results_sim150.zip    12 trees over emptysim.4.100.50 (150 SSMs, no CNVs)
  results_vectors.json  {archive: {option set: the three outputs as JSON values, or {"error": message}}}
The archives are written by this repository's pickler (tests/results_util.py); every number in the vectors is the reference's."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_py3                                                   # noqa: E402
from tests import results_util as U                                          # noqa: E402
from tests.test_write_results import OPTION_SETS                             # noqa: E402

empty = os.path.join(HERE, "_no_cnvs.txt")
open(empty, "w").close()
jobs = {"results_bundled.zip": (os.path.join(HERE, "ssm_data.txt"), os.path.join(HERE, "cnv_data.txt"), 12, 3, {"mh_itr": 5000, "burnin": 4}),
        "results_sim150.zip": (os.path.join(HERE, "emptysim.4.100.50.ssm.txt"), empty, 150, 7, None)}
vectors = {}
for name, (ssm, cnv, n, seed, params) in jobs.items():
    trees, mapping = U.shaped_forest(ssm, cnv, n, seed)
    path = os.path.join(HERE, name)
    U.write_archive(path, trees, mapping, params=params, burnin=trees[:2])
    vectors[name] = {}
    for key, kw in OPTION_SETS.items():
        try:
            vectors[name][key] = ref_py3.reference_results(path, **kw)
        except Exception as e:                                               # noqa: BLE001 - the reference's own refusal is the vector
            vectors[name][key] = {"error": str(e)}
        print(name, key, sorted(vectors[name][key].get("summaries", {}).get("trees", vectors[name][key])))
os.remove(empty)
with open(os.path.join(HERE, "results_vectors.json"), "w") as f:
    json.dump(vectors, f, sort_keys=True)
