"""CPU: the N>1 host logic (read sharding, stats all-reduce, record concatenation) with world_size 2 over gloo.
Each rank fills its shard with the oracle (checker) so the merged result can be compared with a single run."""
import os
import socket
import sys

import numpy as np
import pytest

from biokanga_b200 import abi, sharding

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def test_shard_ranges_cover_and_keep_pairs_together():
    for n in (0, 1, 2, 7, 10, 1001, 20_000_000):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                b, e = sharding.shard_range(n, r, world)
                assert b == prev and e >= b
                assert b % 2 == 0 or b == n  # (an empty trailing shard starts at n)
                prev = e
            assert prev == n


def _worker(rank, world, port, golden_root, out_dir):
    sys.path[:0] = [os.path.dirname(golden_root), os.path.join(os.path.dirname(golden_root), "..", "oracle"), golden_root + "/.."]
    import torch.distributed as dist
    import goldutil as gu
    import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oidx = po.OracleIndex(gu.sfx_path("tiny", out_dir))
    run = gu.runs("tiny")["pe_U2"]
    p, pe = gu.params_from_args(oidx, run["args"])
    names, bases, offs = gu.load_reads("tiny", run)
    n = len(names)
    b, e = sharding.shard_range(n, rank, world)
    local, st = oidx.align(p, bases, offs[b:e + 1])
    pst = oidx.pair(p, pe, local, bases, offs[b:e + 1])
    tot = sharding.all_reduce_stats(st, dist)
    ptot = sharding.all_reduce_stats(pst, dist)
    merged = sharding.gather_results(local, n, rank, world, dist)
    if rank == 0:
        np.save(os.path.join(out_dir, "merged.npy"), merged)
        np.save(os.path.join(out_dir, "stats.npy"), sharding.stats_to_array(tot))
        np.save(os.path.join(out_dir, "pstats.npy"), sharding.stats_to_array(ptot))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_run_equals_single_run(tmp_path):
    import goldutil as gu
    import pyoracle as po
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    gu.sfx_path("tiny", tmp_path)
    mp.spawn(_worker, args=(2, port, gu.GOLD, str(tmp_path)), nprocs=2, join=True)
    oidx = po.OracleIndex(gu.sfx_path("tiny", tmp_path))
    run = gu.runs("tiny")["pe_U2"]
    p, pe = gu.params_from_args(oidx, run["args"])
    names, bases, offs = gu.load_reads("tiny", run)
    exp, st = oidx.align(p, bases, offs)
    pst = oidx.pair(p, pe, exp, bases, offs)
    merged = np.load(tmp_path / "merged.npy")
    assert merged.tobytes() == exp.tobytes()
    assert np.array_equal(np.load(tmp_path / "stats.npy"), sharding.stats_to_array(st))
    assert np.array_equal(np.load(tmp_path / "pstats.npy"), sharding.stats_to_array(pst))
