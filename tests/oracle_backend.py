"""TEST INFRASTRUCTURE: the CPU oracle behind the host samplers' Backend interface, to obtain the reference chain the GPU chain
must reproduce (same host code, same host RNG, same Philox streams => same trajectory)."""
import numpy as np

from oracle import oracle as O
from phylowgs_b200.host.backend import Backend, arrays_from_data


class OracleBackend(Backend):
    def __init__(self, data):
        arr = arrays_from_data(data)
        self.n_ssm = arr["n_ssm"]
        self.data = O.Data(**arr)
        self.tree = None

    def set_tree_arrays(self, parent, node_of_datum, phi, pi):
        self.tree = O.Tree(parent, node_of_datum, phi, pi)

    def llh_rows(self, rows):
        return O.llh_rows(self.data, self.tree, rows, O.SEM_PY)

    def llh_block(self, rows, cols):
        return O.llh_block(self.data, self.tree, rows, cols, O.SEM_PY)

    def total_llh(self):
        return O.total_llh(self.data, self.tree, O.SEM_PY)

    def mh(self, iters, std, seed, stream):
        r = O.mh_loop(self.data, self.tree, iters, std, O.RNG_PHILOX, seed, stream)
        return dict(phi=r["phi"], pi=r["pi"], ratio=r["ratio"], llh=r["llh"])

    def close(self):
        pass
