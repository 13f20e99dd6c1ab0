"""world_size-2 gloo test of the multi-GPU host logic (no GPU): partitioning, the set-up broadcast and
the ordered gather.  On GPUs the same code runs over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from realtime_urdf_filter_b200 import sharding


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
        rng = np.random.default_rng(5)
        arrays = None
        if rank == 0:
            arrays = {"tri": rng.normal(size=(1000, 9)).astype(np.float32),
                      "tri_part": rng.integers(0, 7, 1000).astype(np.int32),
                      "views": rng.normal(size=(6, 16)),
                      "depth": rng.integers(0, 8000, (6, 12, 16)).astype(np.int16),
                      # the library's own buffer types (16UC1 depth, uint32 tri_part) and a 0-d entry in the middle
                      "depth_u16": np.arange(40000, 40096, dtype=np.uint16).reshape(2, 6, 8),
                      "scalar": np.float64(2.5),
                      "part_u32": np.array([0, 3, 0xfffffff0], np.uint32)}
        got = sharding.broadcast_arrays(arrays, 0, "cpu")
        assert list(got) == ["tri", "tri_part", "views", "depth", "depth_u16", "scalar", "part_u32"]
        assert np.array_equal(got["depth_u16"].numpy().view(np.uint16), np.arange(40000, 40096, dtype=np.uint16).reshape(2, 6, 8))
        assert got["scalar"].shape == () and float(got["scalar"]) == 2.5
        assert got["part_u32"].numpy().view(np.uint32).tolist() == [0, 3, 0xfffffff0]
        ref = np.random.default_rng(5)
        assert np.array_equal(got["tri"].numpy(), ref.normal(size=(1000, 9)).astype(np.float32))
        assert got["tri_part"].dtype == torch.int32 and got["views"].dtype == torch.float64
        assert got["depth"].shape == (6, 12, 16)
        n = 11
        mine = sharding.frames_for_rank(n, rank, world)
        local = np.stack([np.full((4,), k, np.int32) for k in mine])     # "result" of frame k
        out = sharding.gather_in_order(local, mine, n)
        if rank == 0:
            assert out[:, 0].tolist() == list(range(n))
        q.put((rank, mine))
    finally:
        dist.destroy_process_group()


def test_two_rank_broadcast_partition_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(world))
    assert sorted(res[0] + res[1]) == list(range(11)) and not set(res[0]) & set(res[1])


def test_partition_helpers():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += sharding.frames_for_rank(37, r, world)
        assert sorted(seen) == list(range(37))
    assert sharding.streams_for_rank(8, 3, 8) == [3]
    assert sharding.streams_for_rank(8, 1, 2) == [1, 3, 5, 7]
