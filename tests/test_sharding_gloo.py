"""CPU, world_size 2, gloo: the host-side logic of the N>1 path (stream sharding, vocabulary-blob broadcast,
counter gather).  The data path itself has no collective."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orb_slam2_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.streams_for_rank(8, rank, world)
        blob = torch.arange(0, 100003, dtype=torch.int64).to(torch.uint8) if rank == 0 else None
        got = sharding.broadcast_blob(blob, src=0)
        counters = sharding.gather_counters([len(mine) * 10, 2000 * len(mine), 1990 * len(mine), 1100 * len(mine), 1234 + rank])
        q.put((rank, mine, int(got.sum()), got.numel(), counters.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5, 7]
    want_sum = int(torch.arange(0, 100003, dtype=torch.int64).to(torch.uint8).sum())
    assert res[0][2] == res[1][2] == want_sum and res[0][3] == res[1][3] == 100003
    assert res[0][4] == res[1][4]
    c = np.array(res[0][4])
    assert c.shape == (2, len(sharding.COUNTER_FIELDS)) and c[:, 0].sum() == 80 and c[1, 4] == 1235


def test_stream_partition_is_a_partition():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 8, 64):
            parts = [sharding.streams_for_rank(n, r, world) for r in range(world)]
            assert sorted(s for p in parts for s in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_random_vocabulary_shape():
    parent, leaf, desc, weight = sharding.random_vocabulary_arrays(10, 3, 1)
    assert len(parent) == 1111 and leaf.sum() == 1000 and desc.shape == (1111, 32)
    assert np.all(parent[1:] < np.arange(1, 1111)) and np.all(weight[leaf == 1] > 0) and np.all(weight[leaf == 0] == 0)
    assert np.bincount(parent[1:]).max() == 10
