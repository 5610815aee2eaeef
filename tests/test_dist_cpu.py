"""CPU, world_size 2, gloo: sharding rule + the single all-gather of skeleton records (host logic of SURVEY 8(e))."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smap_b200 import dist as sdist
from smap_b200._lib import RECORD_BYTES


def test_shard_ranges_cover_and_are_contiguous():
    for n in (1, 7, 8, 64, 65):
        for world in (1, 2, 4, 8):
            lo_prev = 0
            for r in range(world):
                lo, hi = sdist.shard_range(n, r, world)
                assert lo == lo_prev and hi >= lo
                lo_prev = hi
            assert lo_prev == n
            assert max(sdist.shard_sizes(n, world)) - min(sdist.shard_sizes(n, world)) <= 1


def _fake_records(n):
    rng = np.random.default_rng(123)
    return torch.from_numpy(rng.integers(0, 256, (n, RECORD_BYTES), dtype=np.uint8))


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    allrec = _fake_records(n_frames)
    lo, hi = sdist.shard_range(n_frames, rank, world)
    got = sdist.allgather_records(allrec[lo:hi].clone(), n_frames=n_frames)
    q.put((rank, bool(torch.equal(got, allrec))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [8, 7])
def test_allgather_records_world2_gloo(n_frames):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def _tile_worker(rank, world, port, q):
    import torch.distributed as dist

    from smap_b200 import _lib, dist as sdist, engine

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lib = _lib.load()
    # rank 0 "autotuned" two geometries, rank 1 a different choice for one of them: after the sync everyone holds rank 0's
    lib.smapb_set_tile_table(b"G1\t128\t2\nG2\t64\t1\n" if rank == 0 else b"G1\t256\t1\n")
    sdist.sync_tile_table()
    q.put((rank, engine.get_tile_table()))
    dist.destroy_process_group()


def test_sync_tile_table_broadcasts_rank0_choices():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 300
    ps = [ctx.Process(target=_tile_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(60)
    for r in range(2):
        rows = dict(l.split("\t", 1) for l in got[r].strip().split("\n"))
        assert rows["G1"] == "128\t2" and rows["G2"] == "64\t1"
