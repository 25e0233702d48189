"""-m "not gpu": the multi-chain / multi-GPU host logic.  Chains share nothing, so N > 1 means N independent chains: the tests
cover chain -> GPU placement, chain merging (multievolve.py:215-313), and - with a world_size-2 gloo group on CPU - the only
cross-rank step bench.py has: the max-over-ranks reduction of the timings and rank 0 being the reporter."""
import os
import sys
import zipfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chain_to_gpu_placement_is_round_robin():
    from phylowgs_b200.host.multievolve import chain_device
    assert [chain_device(i, list(range(8))) for i in range(8)] == list(range(8))
    assert [chain_device(i, [0, 1]) for i in range(5)] == [0, 1, 0, 1, 0]
    assert [chain_device(i, [3]) for i in range(3)] == [3, 3, 3]


def test_independent_streams_per_rank():
    import bench
    ids = {bench.chain_stream(r, ph, k) for r in range(8) for ph in range(3) for k in range(100)}
    assert len(ids) == 8 * 3 * 100


def _fake_chain(d, llhs, extra=b"{}"):
    os.makedirs(d)
    with zipfile.ZipFile(os.path.join(d, "trees.zip"), "w") as z:
        z.writestr("params.json", extra)
        z.writestr("burnin_-1_-50.0", b"b")
        for i, llh in enumerate(llhs):
            z.writestr("tree_%d_%s" % (i, llh), ("t%d" % i).encode())


def test_chain_selection_and_merge(tmp_path):
    from phylowgs_b200.host import multievolve as M
    dirs = [str(tmp_path / ("chain_%d" % i)) for i in range(3)]
    _fake_chain(dirs[0], [-100.0, -101.0])
    _fake_chain(dirs[1], [-105.0, -104.0, -103.0])
    _fake_chain(dirs[2], [-200.0])
    inc, exc = M.determine_chains_to_merge(dirs, 1.1)
    assert [i for i, _ in inc] == [0, 1] and [i for i, _ in exc] == [2]
    assert inc[0][1] == pytest.approx(np.logaddexp(-100.0, -101.0))
    M.merge_best_chains(str(tmp_path), dirs, inc, exc)
    with zipfile.ZipFile(str(tmp_path / "trees.zip")) as z:
        names = z.namelist()
        assert sorted(n for n in names if n.startswith("tree")) == ["tree_0_-100.0", "tree_1_-101.0", "tree_2_-105.0", "tree_3_-104.0", "tree_4_-103.0"]
        assert "params.json" in names and not any(n.startswith("burnin") for n in names)
    inc_all, exc_all = M.determine_chains_to_merge(dirs, float("inf"))
    assert len(inc_all) == 3 and not exc_all


def _gloo_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    slow = [12.5, 40.25][rank]                               # per-rank elapsed ms
    got = bench.reduce_max_over_ranks(slow, world, torch.device("cpu"))
    dist.barrier()
    with open(os.path.join(out_dir, "r%d.txt" % rank), "w") as f:
        f.write(repr(got))
    dist.destroy_process_group()


def test_max_over_ranks_with_gloo_world_size_2(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    vals = [float(open(str(tmp_path / ("r%d.txt" % r))).read()) for r in range(2)]
    assert vals == [40.25, 40.25]                            # every rank sees the slowest chain's time; rank 0 reports it
