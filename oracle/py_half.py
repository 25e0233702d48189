"""TEST / BASELINE INFRASTRUCTURE (never imported by the product): a plain-Python restatement of the PYTHON half of the
reference's hot path, with the reference's own cost structure, so that `bench.py` can time "what the reference pays per tree
sample on one host core" next to the GPU figure ("Python half").  It runs on the GPU box, where
/root/reference does not exist; in the build container tests/test_py_half.py holds it against the reference's own code
(oracle/ref_py3.py) through the frozen vectors of tests/golden/py_vectors.json.

What is restated, and the property that makes the reference slow, kept on purpose:

* `logprob(tssb, datum, node)` = alleles.logprob (alleles.py:52-53) -> Datum._log_likelihood (data.py:26-29): for EVERY sample
  tp it first re-derives the tree metadata -- set_node_height, set_path_from_root_to_node, map_datum_to_node
  (data.py:33-38, util2.py:98-115; the last one touches every datum of the data set) -- and only then evaluates
  __log_complete_likelihood__ (data.py:47-62) with compute_n_genomes (data.py:65-126) and find_most_recent_cnv (128-134).
* `sweep_seconds_per_datum` = the per-datum body of TSSB.resample_assignments (tssb.py:82-150): slice sampling over the unit
  interval with one `logprob` per candidate node not yet seen for this datum.
* `complete_data_log_likelihood` (tssb.py:404-410, alleles.py:55-56): one `logprob` per datum again.

The objects are duck-typed on the attribute contract the reference's pickles carry (SURVEY.md 8b "Output compat"):
tssb.root / get_nodes(), node.parent() / children() / get_ancestors() / get_data() / params / pi, datum.a / d / mu_r / mu_v /
cnv [(cnv_datum, c1, c2)] / node / _log_bin_norm_const -- phylowgs_b200.host.model provides them.
"""
import math
import time

import numpy as np


# ---------------------------------------------------------------- util2.py:98-115
def set_node_height(tssb):
    tssb.root["node"].ht = 0

    def descend(root, ht):
        for child in root.children():
            child.ht = ht
            descend(child, ht + 1)
    descend(tssb.root["node"], 1)


def set_path_from_root_to_node(tssb):
    for node in tssb.get_nodes():
        node.path = node.get_ancestors()


def map_datum_to_node(tssb):
    for node in tssb.get_nodes():
        for datum in node.get_data():      # a fresh list of the node's Datum objects per call (node.py:60-64)
            datum.node = node


# ---------------------------------------------------------------- data.py:47-134
def log_binomial_likelihood(x, n, mu):     # util2.py:30-31
    return x * math.log(mu) + (n - x) * math.log(1 - mu)


def logsumexp(xs):                         # util2.py:36-38
    m = max(xs)
    return math.log(sum(math.exp(x - m) for x in xs)) + m


def find_most_recent_cnv(datum, nd):       # data.py:128-134
    for n in nd.get_ancestors()[::-1]:
        for c in datum.cnv:
            if n is c[0].node:
                return c
    return None


def compute_n_genomes(tssb, datum, tp):    # data.py:65-126 (new_state = 0)
    nr = [0.0, 0.0, 0.0, 0.0]
    nv = [0.0, 0.0, 0.0, 0.0]
    ssm_node = datum.node.path[-1]
    for nd in tssb.get_nodes():
        pi = nd.pi[tp]
        mr = find_most_recent_cnv(datum, nd)
        inside = any(a is ssm_node for a in nd.get_ancestors())
        if not inside and not mr:
            for q in range(4):
                nr[q] += pi * 2
        elif inside and not mr:
            for q in range(4):
                nr[q] += pi
                nv[q] += pi
        elif not inside and mr:
            for q in range(4):
                nr[q] += pi * (mr[1] + mr[2])
        else:
            t = mr[1] + mr[2]
            for q in (2, 3):
                nr[q] += pi * max(0, t - 1)
                nv[q] += pi * min(1, t)
            if any(a is ssm_node for a in mr[0].node.get_ancestors()):
                nr[0] += pi * mr[1]
                nv[0] += pi * mr[2]
                nr[1] += pi * mr[2]
                nv[1] += pi * mr[1]
            else:
                for q in (0, 1):
                    nr[q] += pi * max(0, t - 1)
                    nv[q] += pi * min(1, t)
    if len(datum.cnv) == 1 and datum.node is datum.cnv[0][0].node:
        return [(nr[q], nv[q]) for q in range(4)]
    return [(nr[0], nv[0]), (nr[1], nv[1])]


def log_complete_likelihood(tssb, datum, phi, tp):     # data.py:47-62
    lbc = float(np.ravel(datum._log_bin_norm_const)[tp])
    a, d = int(datum.a[tp]), int(datum.d[tp])
    if datum.cnv:
        poss = [x for x in compute_n_genomes(tssb, datum, tp) if x[1] > 0]
        ll = []
        for (nr, nv) in poss:
            mu = (nr * datum.mu_r + nv * (1 - datum.mu_r)) / (nr + nv)
            ll.append(log_binomial_likelihood(a, d, mu) + math.log(1.0 / len(poss)) + lbc)
        if not poss:
            ll.append(math.log(1e-99))
        return logsumexp(ll)
    mu = (1 - phi) * datum.mu_r + phi * datum.mu_v
    return log_binomial_likelihood(a, d, mu) + lbc


def logprob(tssb, datum, node):
    """alleles.logprob([datum]) with the datum sitting in `node` (the caller has moved it there, tssb.py:122-124)."""
    params = np.ravel(node.params)
    total = 0.0
    for tp in range(len(params)):
        set_node_height(tssb)                          # data.py:33-38: before every single evaluation
        set_path_from_root_to_node(tssb)
        map_datum_to_node(tssb)
        total += log_complete_likelihood(tssb, datum, float(params[tp]), tp)
    return total


# ---------------------------------------------------------------- tssb.py:82-150, timed on a bounded sample of the data
def sweep_sample(tssb, rows, rng):
    """The body of resample_assignments for the data in `rows` (in place).  -> (seconds, logprob calls)."""
    eps = np.finfo(np.float64).eps
    calls = 0
    t0 = time.perf_counter()
    for n in rows:
        datum = tssb.data[n]
        llhmap = {}
        ancestors = tssb.assignments[n].get_ancestors()
        current, indices = tssb.root, []
        for anc in ancestors[1:]:
            index = [c["node"] for c in current["children"]].index(anc)
            current = current["children"][index]
            indices.append(index)
        max_u, min_u = 1.0, 0.0
        old_llh = logprob(tssb, datum, tssb.assignments[n])
        calls += 1
        llhmap[id(tssb.assignments[n])] = old_llh
        llh_s = math.log(rng.rand()) + old_llh
        while True:
            new_u = (max_u - min_u) * rng.rand() + min_u
            new_node, new_path = tssb.find_node(new_u)
            if new_node.parent() is None:
                new_node, new_path = new_node.children()[0], [0]
            old_node = tssb.assignments[n]
            old_node.remove_datum(n)
            new_node.add_datum(n)
            tssb.assignments[n] = new_node
            if id(new_node) in llhmap:
                new_llh = llhmap[id(new_node)]
            else:
                new_llh = logprob(tssb, datum, new_node)
                calls += 1
                llhmap[id(new_node)] = new_llh
            if new_llh > llh_s:
                break
            new_node.remove_datum(n)
            old_node.add_datum(n)
            tssb.assignments[n] = old_node
            if abs(max_u - min_u) < eps:
                break
            s1 = "".join("%03d" % i for i in indices)
            s2 = "".join("%03d" % i for i in new_path)
            if not indices and not new_path:
                comp = 0
            elif not indices:
                comp = 1
            elif not new_path:
                comp = -1
            else:
                comp = (s2 > s1) - (s2 < s1)
            if comp < 0:
                min_u = new_u
            else:
                max_u = new_u
    return time.perf_counter() - t0, calls


def llh_trace_sample(tssb, rows):
    """The data term of complete_data_log_likelihood (tssb.py:404-410) for the data in `rows`.  -> seconds."""
    t0 = time.perf_counter()
    total = 0.0
    for n in rows:
        total += logprob(tssb, tssb.data[n], tssb.assignments[n])
    return time.perf_counter() - t0, total
