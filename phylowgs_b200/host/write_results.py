"""Summaries of a finished run for the result viewers: the reference's `write_results.py` (write_results.py:17-55) with its
`pwgsresults` package (result_generator.py, result_munger.py, index_calculator.py, json_writer.py), SURVEY 8f-4 -- the last
consumer of trees.zip on the path.  Same command line, same three output files:

    python -m phylowgs_b200.host.write_results [--include-ssm-names] [--min-ssms X] [--include-multiprimary] [--keep-superclones]
        [--max-multiprimary P] dataset_name trees.zip tree_summary.json.gz mutlist.json.gz mutass.zip

  tree_summary  {"dataset_name", "params", "trees": {idx: {"llh", "structure", "populations", three indices}}, "tree_densities"}
  mutlist       {"ssms": {id: reads, error rates[, name]}, "cnvs": {id: reads, physical_cnvs, ssms}, "dataset_name"}
  mutass        one "<idx>.json" per tree: {"mut_assignments": {population: {"ssms": [...], "cnvs": [...]}}, "dataset_name"}

There is no likelihood work here, only bookkeeping on the sampled trees, so this is host Python like the reference's; it is
organised around one small record per tree (`TreeSummary`) instead of the reference's three parallel dictionaries.  The
numbers are the reference's: tests/test_write_results.py holds every output against the reference's own classes executed
under Python 3 (oracle/ref_py3.py) on the same archive, and against vectors frozen from them (tests/golden/results_*.json)."""
import argparse
import gzip
import json
import sys
import zipfile

import numpy as np

from .post_assign_ssm import remove_empty_nodes


# ------------------------------------------------------------------ reading trees.zip (util2.py:264-343)
class Archive(object):
    """Sampled trees of a trees.zip by index, with the LLH each member carries in its name (util2.py:250-262)."""

    def __init__(self, path):
        from .evolve import install_pickle_aliases
        install_pickle_aliases()
        self.zip = zipfile.ZipFile(path)
        members = []
        for info in self.zip.infolist():
            if info.filename.startswith("tree_"):
                tok = info.filename.split("_")
                members.append((int(tok[1]), float(tok[2]), info))
        members.sort(key=lambda m: m[0])
        if not members:
            raise ValueError("%s holds no sampled trees" % path)
        if [m[0] for m in members] != list(range(len(members))):
            raise ValueError("tree indices of %s are not 0..n-1" % path)           # util2.py:279
        self.members = members

    def extra(self, name):
        return self.zip.read(name)

    def by_llh(self):
        """Members from the best tree down (util2.py:335; ties keep index order, the sort is stable)."""
        return sorted(self.members, key=lambda m: m[1], reverse=True)

    def tree(self, info, remove_empty_vertices=False):
        import pickle
        tree = pickle.loads(self.zip.read(info))
        if remove_empty_vertices:
            remove_empty_nodes(tree.root)
        return tree

    def close(self):
        self.zip.close()


# ------------------------------------------------------------------ one tree's summary (result_generator.py:43-93)
class TreeSummary(object):
    """Populations in preorder with children visited by decreasing mean phi (the numbering printo_latex.py uses,
    result_generator.py:75-84): populations[i] = cellular prevalence per sample and mutation counts, structure[i] = children of i
    (only populations that have any), mutass[i] = ids of the SSMs / CNVs of population i (only populations that have any)."""

    def __init__(self, llh, populations, structure, mutass):
        self.llh, self.populations, self.structure, self.mutass = llh, populations, structure, mutass

    @classmethod
    def of_tree(cls, tree, llh):
        populations, structure, mutass = {}, {}, {}
        counter = [0]

        def visit(node):
            me = counter[0]
            ssms, cnvs = [], []
            for mut in node.get_data():
                if mut.id.startswith("s"):
                    ssms.append(mut.id)
                elif mut.id.startswith("c"):
                    cnvs.append(mut.id)
                else:
                    raise ValueError("Unknown mutation ID type: %s" % mut.id)
            if ssms or cnvs:
                mutass[me] = {"cnvs": cnvs, "ssms": ssms}
            populations[me] = {"cellular_prevalence": list(node.params), "num_ssms": len(ssms), "num_cnvs": len(cnvs)}
            for child in sorted(node.children(), key=lambda v: np.mean(v.params), reverse=True):
                counter[0] += 1
                structure.setdefault(me, []).append(counter[0])
                visit(child)

        visit(tree.root["node"])
        return cls(llh, populations, structure, mutass)

    def as_json(self):
        return {"llh": self.llh, "structure": self.structure, "populations": self.populations}

    # -------- tree surgery (result_munger.py:133-299)
    def parent_of(self, idx):
        for parent, children in self.structure.items():
            if idx in children:
                return parent
        raise KeyError("population %s has no parent in %s" % (idx, self.structure))

    def mean_phi(self, idx):
        return np.mean(self.populations[idx]["cellular_prevalence"])

    def remove(self, doomed, mutlist, destination):
        """Delete the populations `doomed`: their children move up to the grandparent, the survivors are renumbered compactly (order
        kept), the mutations of the deleted populations go to the population whose phi is closest to the one their reads imply
        (destination "best", result_munger.py:254-277) or to the clonal population 1 ("clonal", 279-284), counts are refreshed."""
        doomed = sorted(set(doomed))
        if not doomed:
            return
        for idx in doomed:
            del self.populations[idx]
        for idx in doomed:                                                          # result_munger.py:231-252
            parent = self.parent_of(idx)
            kids = [c for c in self.structure[parent] if c != idx] + self.structure.pop(idx, [])
            if kids:
                self.structure[parent] = sorted(kids)
            else:
                del self.structure[parent]
        # compact numbering: a survivor moves down by the number of deleted populations below it (result_munger.py:133-180)
        new_index, gone = {}, 0
        doomed_set = set(doomed)
        top = max(self.populations)
        for idx in range(1, top + 1):
            if idx in doomed_set:
                gone += 1
            elif gone:
                new_index[idx] = idx - gone
        self.populations = {new_index.get(i, i): p for i, p in sorted(self.populations.items())}
        self.structure = {new_index.get(i, i): [new_index.get(c, c) for c in kids] for i, kids in sorted(self.structure.items())}
        orphans = [self.mutass.pop(idx) for idx in doomed if idx in self.mutass]
        self.mutass = {new_index.get(i, i): m for i, m in sorted(self.mutass.items())}
        for muts in orphans:
            for kind in ("ssms", "cnvs"):
                for mut_id in muts[kind]:
                    target = 1 if destination == "clonal" else self.closest_population(mutlist[kind][mut_id])
                    self.mutass.setdefault(target, {"cnvs": [], "ssms": []})[kind].append(mut_id)
        for idx, pop in self.populations.items():                                   # result_munger.py:182-191
            if idx != 0:
                held = self.mutass.setdefault(idx, {"cnvs": [], "ssms": []})
                pop["num_ssms"], pop["num_cnvs"] = len(held["ssms"]), len(held["cnvs"])

    def closest_population(self, reads):
        """The cancerous population whose mean phi is nearest to 2 x VAF of the mutation, capped at 1 (one mutated copy of two
        assumed, result_munger.py:257-275); the first one wins a tie."""
        ref, total = np.mean(reads["ref_reads"]), np.mean(reads["total_reads"])
        implied = min(2 * (total - ref) / total, 1.0)
        best, best_delta = None, 1
        for idx in sorted(self.populations):
            delta = abs(self.mean_phi(idx) - implied)
            if delta < best_delta and idx != 0:
                best, best_delta = idx, delta
        return best

    # -------- shape indices (index_calculator.py:5-88)
    def indices(self):
        """(linearity, branching, clustering): the shares of ordered SSM pairs whose populations are ancestor and descendant, sit on
        different branches, or coincide -- pairs of an SSM with itself left out, so the three add up to one."""
        n = {idx: pop["num_ssms"] for idx, pop in self.populations.items()}
        total = sum(n.values())
        below = {}                                                                  # SSMs in the subtree of each population

        def subtree(idx):
            below[idx] = n[idx] + sum(subtree(c) for c in self.structure.get(idx, ()))
            return below[idx]

        subtree(0)
        pairs = total * (total - 1)
        anc_desc = sum(n[idx] * (below[idx] - n[idx]) for idx in below)            # ordered (ancestor, descendant) pairs
        same = sum(v * (v - 1) for v in n.values())
        cousin = total * total - sum(v * v for v in n.values()) - 2 * anc_desc     # ordered pairs on different branches
        linearity = 2 * (float(anc_desc) / pairs)
        branching = float(cousin) / pairs
        clustering = float(same) / pairs
        assert 0.0 <= linearity <= 1.0 and 0.0 <= branching <= 1.0 and 0.0 <= clustering <= 1.0
        assert np.isclose(clustering, 1 - linearity - branching)
        return linearity, branching, clustering


def list_mutations(tree, include_ssm_names, cnv_physical):
    """Read counts and error rates of every mutation, CNVs with the SSMs they contain (result_generator.py:95-136)."""
    ssms, cnvs, inside = {}, {}, {}
    stack = [tree.root]
    while stack:
        entry = stack.pop()
        for mut in entry["node"].get_data():
            if mut.id.startswith("s"):
                ssms[mut.id] = {"ref_reads": mut.a, "total_reads": mut.d, "expected_ref_in_ref": mut.mu_r, "expected_ref_in_variant": mut.mu_v}
                if include_ssm_names:
                    ssms[mut.id]["name"] = mut.name
                for cnv, maternal, paternal in mut.cnv:
                    inside.setdefault(cnv.id, []).append({"ssm_id": mut.id, "maternal_cn": maternal, "paternal_cn": paternal})
            elif mut.id.startswith("c"):
                cnvs[mut.id] = {"ref_reads": mut.a, "total_reads": mut.d, "physical_cnvs": cnv_physical[mut.id]}
            else:
                raise ValueError("Unknown mutation type: %s" % mut.id)
        stack.extend(reversed(entry["children"]))
    for cnv_id, cnv in cnvs.items():
        cnv["ssms"] = inside.get(cnv_id, [])
    return {"ssms": ssms, "cnvs": cnvs}


def generate(tree_file, include_ssm_names=False):
    """-> ({tree index: TreeSummary}, mutlist, params) (result_generator.py:11-41)."""
    archive = Archive(tree_file)
    try:
        ranked = archive.by_llh()
        cnv_physical = json.loads(archive.extra("cnv_logical_physical_mapping.json"))
        try:
            params = json.loads(archive.extra("params.json"))
        except KeyError:                                                            # archives of older runs
            params = {}
        mutlist = list_mutations(archive.tree(ranked[0][2]), include_ssm_names, cnv_physical)
        trees = {}
        for idx, llh, info in ranked:
            trees[idx] = TreeSummary.of_tree(archive.tree(info, remove_empty_vertices=True), llh)
    finally:
        archive.close()
    return trees, mutlist, params


# ------------------------------------------------------------------ filters over all trees (result_munger.py:16-131)
def small_populations(summary, n_ssms_total, min_ssms):
    """Cancerous populations with fewer SSMs than `min_ssms` (a count if >= 1, else a fraction of all SSMs; result_munger.py:202-229)."""
    threshold = int(min_ssms) if min_ssms >= 1 else int(round(float(min_ssms) * n_ssms_total))
    return [idx for idx, pop in sorted(summary.populations.items()) if idx != 0 and pop["num_ssms"] < threshold]


def remove_small_nodes(trees, mutlist, min_ssms):
    for idx in list(trees):
        s = trees[idx]
        s.remove(small_populations(s, len(mutlist["ssms"]), min_ssms), mutlist, "best")


def remove_superclones(trees, mutlist, log=None):
    """A clonal population with a single child that holds at least three times its SSMs at a cellular prevalence within 0.1 is an
    artefact of tree sampling: the two are merged (prevalence = SSM-weighted mean; result_munger.py:23-83)."""
    for idx in list(trees):
        s = trees[idx]
        top = s.structure.get(0, [])
        assert len(top) > 0
        if len(top) != 1:                                                           # multiprimary: left alone
            continue
        clonal = top[0]
        assert clonal == 1
        if len(s.structure.get(clonal, [])) != 1:
            continue
        child = s.structure[clonal][0]
        assert child == 2
        sup, sub = s.populations[clonal], s.populations[child]
        if sub["num_ssms"] == 0 or not (float(sup["num_ssms"]) / sub["num_ssms"] <= 0.33):
            continue
        sup_cp, sub_cp = np.array(sup["cellular_prevalence"]), np.array(sub["cellular_prevalence"])
        if not (np.mean(np.abs(sup_cp - sub_cp)) <= 0.1):
            continue
        if log:
            log("%s Superclone: %s actual clone: %s" % (idx, sup, sub))
        sup["cellular_prevalence"] = list((sup_cp * sup["num_ssms"] + sub_cp * sub["num_ssms"]) / (sup["num_ssms"] + sub["num_ssms"]))
        s.remove([child], mutlist, "clonal")
        if log:
            log("%s New clonal node: %s" % (idx, sup))


def remove_multiprimary_trees(trees, max_multiprimary, log=None):
    """Drop the trees whose root has several children; too many of them is an error (result_munger.py:85-109).  The surviving
    trees keep their indices: the reference's renumbering pass (result_munger.py:111-131) looks for the dropped indices among the
    keys it has just deleted them from and therefore never moves anything, and the viewers read the files it writes that way."""
    poly = sorted(idx for idx, s in trees.items() if len(s.structure.get(0, [])) != 1)
    frac = len(poly) / float(len(trees))
    if frac >= max_multiprimary:
        raise Exception("%d%% of trees are multiprimary (%s of %s), so not enough to report good posterior." % (100 * frac, len(poly), len(trees)))
    for idx in poly:
        if log:
            log("%s multiprimary tree at idx=%s" % (idx, idx))
        del trees[idx]
    return trees


# ------------------------------------------------------------------ outputs (json_writer.py:11-78)
def tree_densities(indices):
    """Gaussian kernel density of the trees in the (clustering index, branching / (branching + linearity)) plane, at the trees
    themselves; one-dimensional, then zero, when the sample covariance is singular (json_writer.py:11-41)."""
    import scipy.stats
    tidxs = sorted(indices)
    lin, bra, clu = (np.array([indices[t][i] for t in tidxs]) for i in range(3))
    x, y = clu, bra / (bra + lin + 0.0001)
    singular = (np.linalg.LinAlgError, FloatingPointError, ValueError)
    with np.errstate(invalid="raise"):
        try:
            xy = np.vstack((x, y))
            density = list(scipy.stats.gaussian_kde(xy)(xy))
        except singular:
            try:
                density = list(scipy.stats.gaussian_kde(x)(x))
            except singular:
                density = np.zeros(len(x))
    return dict(zip(tidxs, density))


def _plain(o):
    """numpy scalars and arrays as the plain numbers json.dump of Python 2 wrote for them."""
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, np.generic):
        return o.item()
    raise TypeError("%r is not JSON serializable" % (o,))


def _dump_gz(obj, path):
    with gzip.GzipFile(path, "w") as f:
        f.write(json.dumps(obj, default=_plain).encode("ascii"))


def write_summaries(dataset_name, trees, params, path):
    body, indices = {}, {}
    for idx, s in trees.items():
        indices[idx] = s.indices()
        body[idx] = s.as_json()
        body[idx]["linearity_index"], body[idx]["branching_index"], body[idx]["clustering_index"] = indices[idx]
    _dump_gz({"dataset_name": dataset_name, "params": params, "trees": body, "tree_densities": tree_densities(indices)}, path)


def write_mutlist(dataset_name, mutlist, path):
    mutlist["dataset_name"] = dataset_name
    _dump_gz(mutlist, path)


def write_mutass(dataset_name, trees, path):
    with zipfile.ZipFile(path, "w", compression=zipfile.ZIP_DEFLATED, allowZip64=True) as z:
        for idx, s in trees.items():
            z.writestr("%s.json" % idx, json.dumps({"mut_assignments": s.mutass, "dataset_name": dataset_name}, default=_plain))


def _unit_interval(x):
    x = float(x)
    if not 0 <= x <= 1:
        raise argparse.ArgumentTypeError("%s outside range [0, 1]" % x)
    return x


def create_argparser():
    p = argparse.ArgumentParser(description="Write JSON files describing trees", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--include-ssm-names", dest="include_ssm_names", action="store_true",
                   help="Include SSM names in output (which may be sensitive data)")
    p.add_argument("--min-ssms", dest="min_ssms", type=float, default=0.01, help="Minimum number or percent of SSMs to retain a subclone")
    p.add_argument("--include-multiprimary", dest="include_multiprimary", action="store_true",
                   help="Whether to include multiprimary trees in result")
    p.add_argument("--keep-superclones", dest="keep_superclones", action="store_true",
                   help="Whether to keep superclones, which are often artifacts of tree sampling")
    p.add_argument("--max-multiprimary", dest="max_multiprimary", type=_unit_interval, default=0.8,
                   help="Maximum proportion of trees that may be multiprimary if --include-multiprimary is not given; an exception is "
                        "thrown if the proportion of multiprimary trees exceeds this value")
    p.add_argument("dataset_name", help="Name identifying dataset")
    p.add_argument("tree_file", help="File containing sampled trees")
    p.add_argument("tree_summary_output", help="Output file for JSON-formatted tree summaries")
    p.add_argument("mutlist_output", help="Output file for JSON-formatted list of mutations")
    p.add_argument("mutass_output", help="Output file for JSON-formatted list of SSMs and CNVs assigned to each subclone")
    return p


def summarize(tree_file, include_ssm_names=False, min_ssms=0.01, keep_superclones=False, include_multiprimary=False,
              max_multiprimary=0.8, log=None):
    """The whole pipeline of write_results.py:44-50 -> ({tree index: TreeSummary}, mutlist, params)."""
    trees, mutlist, params = generate(tree_file, include_ssm_names)
    remove_small_nodes(trees, mutlist, min_ssms)
    if not keep_superclones:
        remove_superclones(trees, mutlist, log)
    if not include_multiprimary:
        trees = remove_multiprimary_trees(trees, max_multiprimary, log)
    return trees, mutlist, params


def main(argv=None):
    args = create_argparser().parse_args(argv)
    trees, mutlist, params = summarize(args.tree_file, args.include_ssm_names, args.min_ssms, args.keep_superclones,
                                       args.include_multiprimary, args.max_multiprimary, log=lambda m: sys.stdout.write(m + "\n"))
    write_summaries(args.dataset_name, trees, params, args.tree_summary_output)
    write_mutlist(args.dataset_name, mutlist, args.mutlist_output)
    write_mutass(args.dataset_name, trees, args.mutass_output)
    return 0


if __name__ == "__main__":
    sys.exit(main())
