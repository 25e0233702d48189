"""ref_py3.py — TEST INFRASTRUCTURE ONLY (never imported by the product; tests/test_abi.py checks that).

Executes the reference's own Python-2 sources under Python 3, from where they lie under /root/reference, after a DECLARED list
of mechanical text patches (below) and with the handful of Python-2 builtins the sources rely on put back.  Nothing of the
reference is copied into this repository: the sources are read at run time, patched in memory and exec'd.

What this is for:
  * golden vectors for the Python likelihood convention — `Datum._log_likelihood`, `compute_n_genomes` (data.py:26-126),
    `params.write_data_state` / `write_tree` (params.py:68-171) and the reference's own `TSSB.resample_assignments`
    (tssb.py:82-150) — produced by the reference's code, not by a restatement (tests/golden/make_py_golden.py);
  * the reference chain of the end-to-end statistical comparison: the reference's own evolve.do_mcmc driving the unmodified
    mh.cpp (oracle/_ref/mh.o, GSL shim) through its own text files (scripts/ref_chain.py);
  * the reference's own `util2.TreeReader` / `pwgsresults` held against the trees.zip this repo writes.

It needs /root/reference, so it only runs in the build container; everything it produces that the GPU box needs is committed
as a fixture under tests/golden/.

The patches.  Each is syntax-level; none changes what a line computes under Python 2:
  P1  tabs expanded to 8 columns in every file             (Python 2's own rule for mixed tab/space indentation:
                                                              util2.py:44, node.py:64, data.py:122-126)
  P2  `print "..."` / `print x, y`  ->  print(...)          (data.py:109, params.py:158, util.py:110, evolve.py:267)
  P3  `print >>f, x`                ->  print(x, file=f)    (tssb.py:136, util2.py:300,345)
  P4  `import cPickle [as pickle]`  ->  import pickle [...] (util2.py:3, util.py:3, evolve.py:4)
  P5  `except X, e:`                ->  except X as e:      (util.py:60,76)
  P6  tuple-parameter lambda at util2.py:335               -> lambda t: t[1]
  P7  `from numpy import *` followed by `from builtins import min, max, round`
                                                             (numpy >= 2 star-exports them; the numpy 1.x the reference
                                                              ran on did not, so data.py:97 `max(0, ...)` is the builtin)
  P8  `dtype='string'` -> dtype=str (params.py:62); file modes 'rU' -> 'r' (util2.py:55,74); pickle files opened in binary
      mode (util2.py:199,206,209); `freq.keys()` wrapped in list() (evolve.py:268)
  P9  pwgsresults only (`load_results`): `assert new_idx < sidx` with new_idx None (result_munger.py:290; Python 2 orders None
      before every int, Python 3 raises) -> `assert new_idx is None or new_idx < sidx`; `np.linalg.linalg.LinAlgError`
      (json_writer.py:27,33; the private module is gone from numpy 2) -> `np.linalg.LinAlgError`
Python-2 builtins restored in each module's globals before it runs: map / filter / range / zip returning lists, cmp, reduce,
xrange.  `scipy.misc` (comb, logsumexp — removed from SciPy) is provided as a two-function stand-in module, and printo.py
(tree drawing through ete2; evolve.py imports it and never calls it) as an empty one.

Declared, unavoidable differences from a Python-2 run: dict iteration order (util2.py:94 builds the data list from
`data.keys()`: hash order under Python 2, file order here), `str(float)` (12 significant digits under Python 2, shortest
round-trip repr here: c_tree.txt carries more digits), numpy >= 2 underneath.
"""
import builtins
import functools
import os
import re
import sys
import types

REF = os.environ.get("PWGS_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

MODULES = ["util", "node", "util2", "data", "alleles", "tssb", "params", "printo", "evolve"]
# import cycles of the reference (util2 <-> data <-> tssb) resolve under Python 2 only in one order; the same order here
ORDER = ["util", "tssb", "node", "util2", "data", "alleles", "params", "printo", "evolve"]


def available():
    return os.path.exists(os.path.join(REF, "tssb.py"))


def _py2_print(line):
    m = re.match(r"^(\s*)print\s*>>\s*([^,]+),\s*(.*)$", line)
    if m:
        return "%sprint(%s, file=%s)" % (m.group(1), m.group(3), m.group(2).strip())
    m = re.match(r"^(\s*)print\s+(?![>(=])(.+)$", line)
    if m:
        return "%sprint(%s)" % (m.group(1), m.group(2))
    return line


def patch_source(name, src):
    """Apply P1..P8 to the text of reference file `name`.  Returns the patched text."""
    src = src.expandtabs(8)                                                                       # P1
    out = []
    for line in src.split("\n"):
        stripped = line.strip()
        if stripped.startswith("print ") or stripped.startswith("print>>"):
            # only whole statements on one line; the multi-line `print >> fh, \"\"\"...` of tssb.print_graph parses as an
            # expression and is never executed on this path
            if stripped.count('"""') % 2 == 0 and not stripped.endswith("\\"):
                line = _py2_print(line)                                                           # P2, P3
        line = re.sub(r"^(\s*)import cPickle as pickle\s*$", r"\1import pickle", line)           # P4
        line = re.sub(r"^(\s*)import cPickle\s*$", r"\1import pickle as cPickle", line)
        line = re.sub(r"^(\s*)except (\w+), (\w+):", r"\1except \2 as \3:", line)                 # P5
        line = line.replace("lambda (tidx, llh, zinfo): llh", "lambda t: t[1]")                   # P6
        if re.match(r"^\s*from numpy\s+import \*\s*$", line):                                     # P7
            line = line + "; from builtins import min, max, round"
        line = line.replace("dtype='string'", "dtype=str").replace(",'rU')", ",'r')")             # P8
        if name == "util2.py":
            line = line.replace("open(state_fn, 'w')", "open(state_fn, 'wb')")
            line = line.replace("open(self._last_state_fn) as", "open(self._last_state_fn, 'rb') as")
            line = line.replace("open(self._initial_state_fn) as", "open(self._initial_state_fn, 'rb') as")
        if name == "evolve.py":
            line = line.replace("array(freq.keys(),str)", "array(list(freq.keys()),str)")
        out.append(line)
    return "\n".join(out)


def _cmp(a, b):
    return (a > b) - (a < b)


_PY2_BUILTINS = {
    "map": lambda f, *its: list(builtins.map(f, *its)),
    "filter": lambda f, it: list(builtins.filter(f, it)),
    "range": lambda *a: list(builtins.range(*a)),
    "zip": lambda *its: list(builtins.zip(*its)),
    "xrange": builtins.range,
    "cmp": _cmp,
    "reduce": functools.reduce,
}


def _scipy_misc_stub():
    import numpy as np
    from scipy.special import comb
    m = types.ModuleType("scipy.misc")
    m.comb = comb

    def logsumexp(x, axis=None):
        from scipy.special import logsumexp as lse
        return lse(np.asarray(x), axis=axis)
    m.logsumexp = logsumexp
    return m


class Reference(object):
    """The reference's modules, exec'd.  Attributes: util, node, util2, data, alleles, tssb, params, printo, evolve."""

    def __init__(self, mh_dir=None, names=None):
        """mh_dir: directory holding the compiled `mh.o` that params.metropolis launches (params.py:60-61 looks next to
        params.py itself: the exec'd module's __file__ is set to <mh_dir>/params.py).  Default oracle/_ref."""
        if not available():
            raise RuntimeError("reference sources not found under %s" % REF)
        mh_dir = mh_dir or os.path.join(HERE, "_ref")
        names = list(names or ORDER)
        self._saved = {}
        self.modules = {}
        import scipy
        if "scipy.misc" not in sys.modules or not hasattr(sys.modules["scipy.misc"], "comb"):
            stub = _scipy_misc_stub()
            self._saved["scipy.misc"] = sys.modules.get("scipy.misc")
            sys.modules["scipy.misc"] = stub
            scipy.misc = stub
        for n in names:
            self._saved[n] = sys.modules.get(n)
        pending = {}
        for n in names:
            path = os.path.join(REF, n + ".py")
            if n == "printo":
                # printo.py (ASCII / LaTeX tree drawing through ete2, imported by evolve.py:16 and never called by it) holds
                # Python-2 string escapes ('\\usepackage') and needs ete2: an empty stand-in module takes its place
                src = ""
            else:
                with open(path) as f:
                    src = patch_source(n + ".py", f.read())
            mod = types.ModuleType(n)
            mod.__file__ = os.path.join(mh_dir, n + ".py") if n == "params" else path
            mod.__dict__.update(_PY2_BUILTINS)
            pending[n] = (mod, compile(src, path, "exec"))
            self.modules[n] = mod
        # a lazy finder: `import X` of a reference module that has not run yet runs it first (what Python's own import does)
        ref = self

        class _Finder(object):
            @staticmethod
            def find_spec(fullname, path=None, target=None):
                if fullname in pending and fullname not in sys.modules:
                    import importlib.machinery

                    class _Loader(object):
                        @staticmethod
                        def create_module(spec):
                            return pending[fullname][0]

                        @staticmethod
                        def exec_module(module):
                            exec(pending[fullname][1], module.__dict__)
                    return importlib.machinery.ModuleSpec(fullname, _Loader())
                return None
        finder = _Finder()
        sys.meta_path.insert(0, finder)
        for n in names:                                      # make sure our copies, not stale same-named modules, are found
            sys.modules.pop(n, None)
        try:
            import importlib
            for n in names:
                importlib.import_module(n)
        finally:
            sys.meta_path.remove(finder)
        for n in names:
            setattr(self, n, sys.modules[n])
            self.modules[n] = sys.modules[n]
        self.names = names

    def release(self):
        """Take the reference's modules out of sys.modules again (they share names with nothing in this repo, but tests that
        install pickle aliases `tssb`, `alleles`, `node`, `data` must not find these instead)."""
        for n in self.names:
            if sys.modules.get(n) is self.modules[n]:
                del sys.modules[n]
            if self._saved.get(n) is not None:
                sys.modules[n] = self._saved[n]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()


def load(**kw):
    return Reference(**kw)


RESULT_MODULES = ["index_calculator", "result_generator", "result_munger", "json_writer"]


def load_results(ref):
    """The reference's pwgsresults package (write_results.py:3-5), exec'd like the rest (P1-P9), on top of a loaded `Reference`
    (result_generator.py imports the reference's util2).  -> namespace with ResultGenerator, ResultMunger, IndexCalculator,
    calc_tree_densities.  JsonWriter's file writing is Python-2 only (json.dump into a binary gzip stream); callers compare the
    structures it would dump."""
    assert "util2" in ref.names
    pkg = types.ModuleType("pwgsresults")
    pkg.__path__ = [os.path.join(REF, "pwgsresults")]
    saved = {n: sys.modules.get(n) for n in ["pwgsresults"] + ["pwgsresults." + m for m in RESULT_MODULES]}
    sys.modules["pwgsresults"] = pkg
    out = types.SimpleNamespace()
    try:
        for name in RESULT_MODULES:
            path = os.path.join(REF, "pwgsresults", name + ".py")
            with open(path) as f:
                src = patch_source(name + ".py", f.read())
            src = src.replace("assert new_idx < sidx", "assert new_idx is None or new_idx < sidx")      # P9
            src = src.replace("np.linalg.linalg.LinAlgError", "np.linalg.LinAlgError")
            mod = types.ModuleType("pwgsresults." + name)
            mod.__file__ = path
            sys.modules["pwgsresults." + name] = mod
            setattr(pkg, name, mod)
            exec(compile(src, path, "exec"), mod.__dict__)
        out.ResultGenerator = sys.modules["pwgsresults.result_generator"].ResultGenerator
        out.ResultMunger = sys.modules["pwgsresults.result_munger"].ResultMunger
        out.IndexCalculator = sys.modules["pwgsresults.index_calculator"].IndexCalculator
        out.calc_tree_densities = sys.modules["pwgsresults.json_writer"].calc_tree_densities
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    return out


def reference_results(tree_file, include_ssm_names=False, min_ssms=0.01, keep_superclones=False, include_multiprimary=False,
                      max_multiprimary=0.8):
    """What the reference's write_results.py (lines 44-55) would put into its three output files for `tree_file`, as plain JSON
    values: {"summaries": tree_summary_output, "mutlist": mutlist_output, "mutass": {tree index: member of mutass_output}}
    (dataset_name left out).  Every number is computed by the reference's own classes."""
    import contextlib
    import io
    import json
    ref = load()
    try:
        R = load_results(ref)
        with contextlib.redirect_stdout(io.StringIO()):
            summaries, mutlist, mutass, params = R.ResultGenerator().generate(tree_file, include_ssm_names)
            munger = R.ResultMunger(summaries, mutlist, mutass)
            summaries, mutass = munger.remove_small_nodes(min_ssms)
            if not keep_superclones:
                munger.remove_superclones()
            if not include_multiprimary:
                munger.remove_multiprimary_trees(max_multiprimary)
        for summary in summaries.values():                                   # json_writer.py:55-59
            calc = R.IndexCalculator(summary)
            summary["linearity_index"] = calc.calc_linearity_index()
            summary["branching_index"] = calc.calc_branching_index()
            summary["clustering_index"] = calc.calc_clustering_index()
        out = {"summaries": {"params": params, "trees": summaries, "tree_densities": R.calc_tree_densities(summaries)},
               "mutlist": mutlist, "mutass": {idx: {"mut_assignments": m} for idx, m in mutass.items()}}

        def plain(o):
            import numpy as np
            if isinstance(o, np.ndarray):
                return o.tolist()
            if isinstance(o, np.generic):
                return o.item()
            raise TypeError(repr(o))
        return json.loads(json.dumps(out, default=plain))
    finally:
        ref.release()
