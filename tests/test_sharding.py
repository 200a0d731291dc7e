"""CPU: the N>1 host logic (stream partition + result gather) with world_size-2 gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from keyword_spotting_b200 import sharding


def test_partition_covers_every_stream_once():
    for total in (0, 1, 7, 131072, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [sharding.partition(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + c0 == s1
            assert blocks[-1][0] + blocks[-1][1] == total
            counts = [c for _, c in blocks]
            assert max(counts) - min(counts) <= 1
            for sid in {0, total // 3, total - 1} - {-1}:
                if 0 <= sid < total:
                    r = sharding.owner_of(sid, total, world)
                    s, c = blocks[r]
                    assert s <= sid < s + c
    with pytest.raises(ValueError):
        sharding.partition(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, count = sharding.partition(total, world, rank)
        local = torch.arange(start, start + count, dtype=torch.int32).unsqueeze(1).repeat(1, 2)
        local[:, 1] = rank
        full = sharding.gather_stream_results(local, total)
        tmax = sharding.reduce_max_scalar(10.0 + rank)
        tsum = sharding.reduce_sum_scalar(float(count))
        q.put((rank, full.numpy(), tmax, tsum))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_stream_results_gloo_world2():
    world, total = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, full, tmax, tsum in results:
        assert full.shape == (total, 2)
        np.testing.assert_array_equal(full[:, 0], np.arange(total))
        np.testing.assert_array_equal(full[:, 1], [sharding.owner_of(i, total, world) for i in range(total)])
        assert tmax == 11.0 and tsum == float(total)


def test_single_process_passthrough():
    x = torch.arange(5)
    assert sharding.gather_stream_results(x, 5) is x
    assert sharding.reduce_max_scalar(3.5) == 3.5
