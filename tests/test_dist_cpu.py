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
