"""The MH kernel's work partition (host code, no device): whatever the shape, every CNV-affected SSM gets exactly one slot, every
class's plain ranges tile the plain list, everything is dealt in whole warp rounds, and the loads are level."""
import ctypes as C

import numpy as np
import pytest

from phylowgs_b200 import _lib

W = np.array([100, 183, 429, 675], dtype=np.int32)
NCLS_MAX = 6                                    # MH_PCLASSES: a warp's 1st .. 6th proposal group of a 256-proposal batch


def plan(cpc, n_states, Dp, ncls, w=W):
    L = _lib.load()
    M = int(sum(n_states))
    ns = np.array(n_states, dtype=np.int32)
    round_slot = np.full(max((M + 31) // 32, 1), -1, dtype=np.int32)
    cta_nc = np.zeros(cpc, dtype=np.int32)
    pstart = np.zeros((NCLS_MAX, cpc + 1), dtype=np.int32)
    cc, pc = C.c_int32(0), C.c_int32(0)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))      # noqa: E731
    _lib.check(L.pwgs_plan_partition(cpc, M, p(ns), Dp, ncls, p(np.ascontiguousarray(w)), p(round_slot), p(cta_nc), p(pstart),
                                     C.byref(cc), C.byref(pc)))
    return round_slot[:(M + 31) // 32], cta_nc, pstart, cc.value, pc.value


SHAPES = [(148, (0, 5732, 9183), 35385), (148, (0, 0, 0), 20000), (148, (7, 0, 0), 0), (148, (100, 2000, 31), 33), (40, (3, 500, 4000), 12345),
          (1, (1, 2, 3), 100), (32, (0, 0, 40000), 1), (148, (5000, 5000, 5000), 500000), (3, (0, 0, 0), 0), (148, (0, 31, 1), 31)]


@pytest.mark.parametrize("cpc,n_states,Dp", SHAPES)
@pytest.mark.parametrize("ncls", [1, 2, 3, 6])
def test_partition_covers_everything_once(cpc, n_states, Dp, ncls):
    M = sum(n_states)
    round_slot, cta_nc, pstart, cchunk, pchunk = plan(cpc, n_states, Dp, ncls)
    # CNV list: sorted rank r lives in slot round_slot[r // 32] + r % 32; slots are distinct and inside their CTA's filled prefix
    assert cta_nc.sum() == M and (cta_nc >= 0).all() and (cta_nc <= cchunk).all() and cchunk % 32 == 0
    slots = np.array([round_slot[r >> 5] + (r & 31) for r in range(M)], dtype=np.int64)
    assert len(set(slots.tolist())) == M
    if M:
        cta, local = slots // max(cchunk, 1), slots % max(cchunk, 1)
        for c in range(cpc):
            mine = np.sort(local[cta == c])
            assert len(mine) == cta_nc[c] and (mine == np.arange(cta_nc[c])).all()          # a dense prefix of the CTA's chunk
        # most expensive first inside a CTA: local order follows sorted rank
        for c in range(cpc):
            r_of = np.nonzero(cta == c)[0]
            assert (np.diff(local[r_of]) > 0).all()
    # plain list: per class, the CTAs' ranges tile [0, Dp) in whole rounds; unused classes repeat class 0
    for i in range(NCLS_MAX):
        b = pstart[i]
        assert b[0] == 0 and b[-1] == Dp and (np.diff(b) >= 0).all()
        assert all(x % 32 == 0 or x == Dp for x in b)
        if i >= ncls:
            assert (b == pstart[0]).all()
    lo, hi = pstart[:, :-1].min(axis=0), pstart[:, 1:].max(axis=0)
    assert (hi - lo).max() <= pchunk
    assert ((pstart[:, 1:] - pstart[:, :-1]).max(axis=0) - (pstart[:, 1:] - pstart[:, :-1]).min(axis=0)).max() <= 32 or Dp % 32


def test_partition_levels_the_cost():
    """config 3's shape: loads (in weight units) within one third of a plain round of each other, but for CTAs that hold nothing
    but expensive rounds."""
    cpc, n_states, Dp = 148, (0, 5732, 9183), 35385
    round_slot, cta_nc, pstart, cchunk, pchunk = plan(cpc, n_states, Dp, 1)
    M = sum(n_states)
    load = np.zeros(cpc)
    for u in range((M + 31) // 32):
        r = 32 * u
        cls = 3 if r < n_states[0] else (2 if r < n_states[0] + n_states[1] else 1)
        load[round_slot[u] // cchunk] += W[cls]
    for ncls in (3, 6):
        round_slot, cta_nc, pstart, cchunk, pchunk = plan(cpc, n_states, Dp, ncls)
        frac = sum(np.ceil(np.diff(pstart[i]) / 32.0) for i in range(ncls))
        total = load + frac * W[0] / ncls
        assert total.max() - total.min() <= W[0] / ncls + 1e-9
        assert total.max() / total.mean() < 1.02
    return
    assert load.max() / load.mean() < 1.02
