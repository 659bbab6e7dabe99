"""Multi-GPU host logic on CPU: world_size-2 gloo run of the shard partition and the trajectory
all-gather (the path's only collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from polychase_b200 import shard


def test_partition_covers_every_pair_once():
    first, num = 5, 103
    for world in (1, 2, 3, 4, 8):
        seen = []
        total = 0
        for r in range(world):
            s, c = shard.shard_range(first, num, world, r)
            total += c
            seen += shard.owned_pairs(first, num, s, c)
            assert shard.halo_frames(first, s) == min(8, s - first)
        assert total == num
        assert len(seen) == len(set(seen)) == 8 * num - 30          # 8F-30 pairs
        want = {(a, a + d) for a in range(first, first + num) for d in (-8, -4, -2, -1, 1, 2, 4, 8)
                if first <= a + d < first + num}
        assert set(seen) == want


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
    first, num = 1, 37
    counts = [shard.shard_range(first, num, world, r)[1] for r in range(world)]
    s, c = shard.shard_range(first, num, world, rank)
    local = torch.zeros((c, shard.CAMERA_STATE_FLOATS))
    local[:, 0] = torch.arange(s, s + c, dtype=torch.float32)       # frame id in slot 0
    local[:, 15] = 1.0                                              # filled
    full = shard.allgather_trajectory(local, counts)
    ok = full.shape == (num, 16) and torch.equal(full[:, 0], torch.arange(first, first + num, dtype=torch.float32))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_allgather_trajectory_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
