"""CPU tests of the batch-sharding host logic, including a world_size-2 gloo run."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import PKG
import sharded


def test_shard_ranges_partition_the_batch():
    for total, world in [(4096, 8), (4096, 1), (10, 4), (3, 8), (0, 2)]:
        ranges = [sharded.shard_range(total, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
            assert a1 == b0 and a1 >= a0
        sizes = sharded.shard_sizes(total, world)
        assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
    assert sharded.shard_sizes(4096, 8) == [512] * 8
    with pytest.raises(ValueError):
        sharded.shard_range(8, 2, 2)


class _FakeNet:
    """Stands in for _lib.Net on a machine without GPUs: out[i] = f(image i) only."""

    def __init__(self, device, batch):
        self.device, self.batch = device, batch
        self.out_shape = (batch, 2)
        self.pending = None

    def forward_async(self, x, out):
        self.pending = (x, out)

    def sync(self):
        x, out = self.pending
        out[:, 0] = x.reshape(x.shape[0], -1).sum(1)
        out[:, 1] = self.device


def test_multi_gpu_engine_covers_every_image_once():
    x = np.random.default_rng(0).random((11, 4, 4, 3), dtype=np.float32)
    eng = sharded.MultiGpuEngine(_FakeNet, 11, devices=[0, 1, 2, 3])
    out = eng.run(x)
    np.testing.assert_allclose(out[:, 0], x.reshape(11, -1).sum(1), rtol=1e-6)
    assert out[:, 1].tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3]


def _worker(rank, world, port, total, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, PKG)
    import sharded as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, stop = sh.shard_range(total, world, rank)
    images = torch.arange(total, dtype=torch.float32).reshape(total, 1, 1, 1).expand(total, 2, 2, 3)
    local = images[start:stop].sum(dim=(1, 2, 3)).reshape(-1, 1) * torch.ones(1, 5)   # "forward" of the shard
    full = sh.gather_outputs(local.contiguous(), total, dist, dst=0)
    if rank == 0:
        q.put(full.numpy())
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_world_size_2_gloo_gather_equals_single_process(total):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    expect = (np.arange(total, dtype=np.float32) * 12).reshape(-1, 1) * np.ones((1, 5), np.float32)
    np.testing.assert_array_equal(full, expect)


def _gather_worker(rank, world, port, n_local, corrupt, q):
    """The package's OverlappedGather + check_shards, as bench.py drives them, on CPU tensors over gloo."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, PKG)
    import sharded as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = lambda idx: torch.stack([torch.full((3, 2), float(i) * 7.0 + 0.25) for i in idx])    # "network": image index -> output
    gat = sh.OverlappedGather((n_local, 3, 2), dist, device=None)
    start = rank * n_local
    for step in range(3):                                   # three steps through the double buffer
        buf = gat.begin()
        buf.copy_(f(range(start, start + n_local)))
        if corrupt and rank == 1 and step == 2:
            buf[n_local - 1, 0, 0] += 1.0                   # a wrong value in the last image of rank 1's slice
        gat.end()
    gat.drain()
    pos = sh.sample_positions(n_local)
    refs = f([start + p for p in pos])                      # the same images "through a batch-1 network"
    res = sh.check_shards(gat.last(), refs, pos, dist, gat.gathered)
    if rank == 0:
        q.put((res, torch.cat(gat.gathered, 0).numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("corrupt", [False, True])
def test_overlapped_gather_and_shard_check_over_gloo(corrupt):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, 5, corrupt, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res["ranks"] == 2 and res["positions_per_rank"] == [0, 2, 4]
    if corrupt:
        assert not res["bit_identical"] and res["mismatches"] >= 1
    else:
        assert res["bit_identical"] and res["mismatches"] == 0
        np.testing.assert_array_equal(full[:, 0, 0], np.arange(10, dtype=np.float32) * 7 + 0.25)    # rank r's slice in slot r


def test_multi_gpu_engine_rejects_buffers_it_would_overrun():
    eng = sharded.MultiGpuEngine(_FakeNet, 4, devices=[0, 1])
    with pytest.raises(ValueError):
        eng.run(np.zeros((4, 4, 4, 3), np.float64))
    with pytest.raises(ValueError):
        eng.run(np.zeros((8, 4, 4, 3), np.float32)[::2])
    with pytest.raises(ValueError):
        eng.run(np.zeros((4, 4, 4, 3), np.float32), np.zeros((3, 2), np.float32))
