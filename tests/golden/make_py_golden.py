"""Generates tests/golden/py_vectors.json from the reference's OWN Python code, executed under Python 3 by oracle/ref_py3.py
(declared mechanical patches only; nothing restated).  Run in the build container, where /root/reference exists:

    python tests/golden/make_py_golden.py

Frozen per case (inputs are the seeded cases of tests/cases.py, the bundled example, and tests/refgold.edge_case()):
  * `grid[j][k]`       = `node_k.logprob([datum_j])` with datum j moved into node k, exactly as TSSB.resample_assignments asks
                         for it (tssb.py:122-128 -> alleles.py:52-53 -> Datum._log_likelihood, data.py:26-62), summed over samples;
  * `n_genomes[j][k]`  = `datum_j.compute_n_genomes(tp)` per sample for the same placement (data.py:65-126), CNV-affected SSMs only
                         (a seeded subset of them in the larger cases, to keep the fixture small);
  * `data_term`        = sum over nodes of `node.data_log_likelihood()` (alleles.py:55-56) for the case's own assignment — the
                         data part of TSSB.complete_data_log_likelihood (tssb.py:404-410);
  * `cdll`             = TSSB.complete_data_log_likelihood() itself (with the seeded sticks of tests/refgold.sticks_for);
  * `data_states`      = the lines of the file params.write_data_state writes (params.py:114-171), parsed to integers (the first
                         600 verbatim, all of them as a SHA-256 of the sorted integer table);
  * `c_tree`           = the lines params.write_tree writes (params.py:68-102);
  * `sweep`            = the state TSSB.resample_assignments (tssb.py:82-150) leaves behind when its draws (`rand`, `boundbeta`)
                         come from a fixed array of uniforms (tests/refgold.ArrayStream): per-datum node paths, per-node sticks /
                         pi / params, the number of uniforms consumed.
"""
import hashlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O       # noqa: E402
from oracle import ref_py3           # noqa: E402
from tests import cases, refgold     # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "py_vectors.json")

# (name, rows of the grid that are frozen: "all" or "cnv+N" = every CNV-affected SSM, every CNV datum and N seeded plain SSMs)
CASES = [("bundled_cnv", "all"), ("edge", "all"), ("synth_0", "cnv+40"), ("synth_1", "cnv+40"), ("synth_2", "all"),
         ("synth_3", "cnv+60"), ("synth_4", "cnv+40")]
SWEEPS = [("bundled_cnv", 3), ("edge", 5), ("synth_0", 3), ("synth_2", 4), ("synth_4", 1), ("sim_4_100_50", 1)]


def case_arrays(name):
    if name == "edge":
        return refgold.edge_case()
    if name == "sim_4_100_50":
        data = O.read_ssm_cnv(os.path.join(cases.GOLDEN, "emptysim.4.100.50.ssm.txt"))
        rng = np.random.default_rng(77)
        parent, nod, phi, pi = cases.random_tree(rng, 7, data.T, data.D, data.n_ssm)
        tree = O.Tree(parent, nod, phi, pi)
    elif name.startswith("bundled"):
        data, tree = cases.bundled(O, name.endswith("cnv"))
    else:
        data, tree = cases.synthetic_case(O, **cases.GOLDEN_SYNTH[int(name.split("_")[1])])
    return dict(a=data.a, d=data.d, mu_r=data.mu_r, mu_v=data.mu_v, n_ssm=data.n_ssm, cnv_ptr=data.cnv_ptr, cnv_idx=data.cnv_idx,
                cnv_c1=data.cnv_c1, cnv_c2=data.cnv_c2, parent=tree.parent, node_of_datum=tree.node_of_datum, phi=tree.phi, pi=tree.pi)


def digest(arr):
    h = hashlib.sha256()
    for key in ("a", "d", "mu_r", "mu_v", "cnv_ptr", "cnv_idx", "cnv_c1", "cnv_c2", "parent", "node_of_datum", "phi", "pi"):
        h.update(np.ascontiguousarray(arr[key]).tobytes())
    return h.hexdigest()


def grid_rows(arr, spec, seed):
    n_ssm, D = int(arr["n_ssm"]), len(arr["a"])
    if spec == "all":
        return list(range(D))
    n_plain = int(spec.split("+")[1])
    cnv_aff = np.nonzero(np.diff(arr["cnv_ptr"]) > 0)[0].tolist()
    plain = [j for j in range(n_ssm) if j not in set(cnv_aff)]
    rng = np.random.default_rng(seed)
    pick = rng.choice(plain, size=min(n_plain, len(plain)), replace=False).tolist() if plain else []
    return sorted(set(cnv_aff + pick + list(range(n_ssm, D))))


def reference_tree(R, arr, seed):
    data = refgold.make_data(R.data.Datum, arr)
    main, stick = refgold.sticks_for(arr["parent"], seed)
    tssb, nodes = refgold.build_tssb(R.tssb.TSSB, R.alleles.alleles, data, arr["parent"], arr["node_of_datum"], arr["phi"], arr["pi"],
                                     main, stick)
    return tssb, nodes, data


def vectors_for(R, name, spec, seed):
    arr = case_arrays(name)
    tssb, nodes, data = reference_tree(R, arr, seed)
    K, T, D, n_ssm = len(nodes), len(data[0].a), len(data), int(arr["n_ssm"])
    out = dict(name=name, sha256=digest(arr), K=K, T=T, D=D, sticks_seed=seed)
    rows = grid_rows(arr, spec, seed)
    out["rows"] = rows
    grid, ngen = [], {}
    cnv_rows = [j for j in rows if data[j].cnv]
    ng_rows = set(cnv_rows if len(cnv_rows) * K * T <= 800 else
                  np.random.default_rng(seed).choice(cnv_rows, size=max(8, 800 // (K * T)), replace=False).tolist())
    for j in rows:
        dat, home = data[j], tssb.assignments[j]
        line, ng_line = [], []
        for k in range(K):
            # tssb.py:122-128: the datum is moved, then the candidate node is asked
            tssb.assignments[j].remove_datum(j)
            nodes[k].add_datum(j)
            tssb.assignments[j] = nodes[k]
            line.append(float(nodes[k].logprob([dat])))
            if dat.cnv and j in ng_rows:
                # logprob has just refreshed heights / paths / datum.node (data.py:33-38)
                ng_line.append([[[float(nr), float(nv)] for (nr, nv) in dat.compute_n_genomes(tp)] for tp in range(T)])
        tssb.assignments[j].remove_datum(j)
        home.add_datum(j)
        tssb.assignments[j] = home
        grid.append(line)
        if dat.cnv and j in ng_rows:
            ngen[str(j)] = ng_line
    out["grid"] = grid
    out["n_genomes"] = ngen
    # the case's own assignment again (the last logprob left datum.node pointing at the last candidate)
    R.util2.set_node_height(tssb)
    R.util2.set_path_from_root_to_node(tssb)
    R.util2.map_datum_to_node(tssb)
    out["data_term"] = float(sum(nd.data_log_likelihood() for nd in nodes if nd.num_local_data()))
    out["cdll"] = float(tssb.complete_data_log_likelihood())
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "c_data_states.txt")
        R.util2.map_datum_to_node(tssb)
        R.params.write_data_state(tssb, fn)
        states = []
        for ln in open(fn):
            tok = ln.rstrip("\n").split("\t")
            if len(tok) < 5:
                continue
            st = [[int(x) for x in tok[q].split(",")] for q in range(1, 5)]
            assert len(set(q[0] for q in st)) == 1          # one node per line (params.py:167-168)
            states.append([int(tok[0]), st[0][0]] + [x for q in st for x in q[1:]])
        # every line as [ssm, node, nr1, nv1, .., nr4, nv4]: the first 600 verbatim (file order), all of them as a SHA-256
        # of the int32 table sorted by (ssm, node)
        out["data_states"] = states[:600]
        out["data_states_lines"] = len(states)
        out["data_states_sha256"] = hashlib.sha256(np.array(sorted(states), dtype=np.int32).tobytes()).hexdigest()
        fn = os.path.join(td, "c_tree.txt")
        R.params.write_tree(tssb, n_ssm, fn)
        out["c_tree"] = [ln.rstrip("\n") for ln in open(fn)]
    return out


def sweep_for(R, name, seed):
    arr = case_arrays(name)
    tssb, nodes, data = reference_tree(R, arr, seed)
    D = len(data)
    uniforms = np.random.RandomState(9000 + seed).rand(16 * D + 8192)
    stream = refgold.ArrayStream(uniforms)
    saved = (R.tssb.rand, R.tssb.boundbeta, R.alleles.rand)
    # the draws of the sweep: tssb.py:110,114 `rand()`, find_node's `boundbeta(1, .)` (tssb.py:353-358), alleles.py:35 `rand(1)`
    R.tssb.rand = lambda *a: stream.next()
    R.alleles.rand = lambda *a: stream.next()

    def bb(a, b):
        assert a == 1 or a == 1.0
        return refgold.boundbeta1(stream.next(), float(b))
    R.tssb.boundbeta = bb
    err = io.StringIO()
    old_err = sys.stderr
    sys.stderr = err
    try:
        tssb.resample_assignments()
    finally:
        sys.stderr = old_err
        R.tssb.rand, R.tssb.boundbeta, R.alleles.rand = saved
    out = refgold.tree_summary(tssb)
    out.update(name=name, sticks_seed=seed, uniform_seed=9000 + seed, n_uniforms=len(uniforms), uniforms_used=stream.pos,
               shrank=err.getvalue().count("Slice sampler shrank"), sha256=digest(arr))
    return out


def main():
    assert ref_py3.available()
    R = ref_py3.load()
    vec = dict(generator="tests/golden/make_py_golden.py",
               source="the reference's own data.py / params.py / tssb.py / alleles.py / node.py / util2.py, exec'd under Python 3 by "
                      "oracle/ref_py3.py (patches P1-P8 listed there)", cases=[], sweeps=[])
    for i, (name, spec) in enumerate(CASES):
        t0 = time.time()
        vec["cases"].append(vectors_for(R, name, spec, 50 + i))
        print("case", name, "%.1f s" % (time.time() - t0), flush=True)
    for name, seed in SWEEPS:
        t0 = time.time()
        vec["sweeps"].append(sweep_for(R, name, seed))
        s = vec["sweeps"][-1]
        print("sweep", name, "%.1f s" % (time.time() - t0), "uniforms", s["uniforms_used"], "nodes", s["n_nodes"], flush=True)
    with open(OUT, "w") as f:
        json.dump(vec, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
