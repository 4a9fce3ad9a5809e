"""Batch-sharded multi-GPU driver.  Images are independent (inference-mode BatchNorm uses stored
statistics, proj3/dnn_openblas.c:59-60), so a global batch is cut into contiguous slices, one per
GPU, weights are replicated once per GPU, and no collective touches the hot path; the only
exchange is the final gather of the [b,13,13,125] outputs.  The reference
has no multi-device code at all.

Two ways to drive it:
  * one process per GPU under torchrun (bench.py): shard_range() + gather_outputs() over
    torch.distributed (NCCL on the box, gloo in the CPU tests);
  * one process, several GPUs: MultiGpuEngine queues every shard asynchronously
    (dnnb_net_forward_async on each device's own stream) and then waits for all of them.
"""
import numpy as np


def shard_range(total, world_size, rank):
    """Contiguous slice [start, stop) of `total` images for `rank`; the first total % world_size
    ranks get one extra image."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(total, world_size):
    return [shard_range(total, world_size, r)[1] - shard_range(total, world_size, r)[0] for r in range(world_size)]


def gather_outputs(local_out, total, dist, dst=0):
    """Final output gather over torch.distributed: every rank passes its [n_r, ...] tensor; rank
    `dst` gets the [total, ...] result in shard order (None elsewhere).  Works for equal and
    ragged shards."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(total, world)
    assert local_out.shape[0] == sizes[rank], (local_out.shape, sizes, rank)
    if world == 1:
        return local_out
    # gather needs equal shapes: ragged shards are padded to the largest and trimmed after
    big = max(sizes)
    send = local_out
    if local_out.shape[0] != big:
        send = torch.zeros((big,) + tuple(local_out.shape[1:]), dtype=local_out.dtype, device=local_out.device)
        send[: local_out.shape[0]] = local_out
    parts = [torch.empty_like(send) for _ in sizes] if rank == dst else None
    dist.gather(send.contiguous(), parts, dst=dst)
    if rank != dst:
        return None
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)


class MultiGpuEngine:
    """Single-process driver: `make_net(device, batch)` builds one resident network per device
    (dnn-inference-engine_b200/_lib.Net or anything with forward_async/sync/out_shape)."""

    def __init__(self, make_net, total_batch, devices):
        self.devices = list(devices)
        self.total = total_batch
        self.ranges = [shard_range(total_batch, len(self.devices), r) for r in range(len(self.devices))]
        self.nets = [make_net(dev, stop - start) if stop > start else None
                     for dev, (start, stop) in zip(self.devices, self.ranges)]
        self.out_shape = (total_batch,) + tuple(next(n for n in self.nets if n is not None).out_shape[1:])

    def run(self, images, out=None):
        """images: [total, H, W, C] float32 host array (page-locked for real overlap)."""
        assert images.shape[0] == self.total
        if out is None:
            out = np.empty(self.out_shape, dtype=np.float32)
        for net, (start, stop) in zip(self.nets, self.ranges):
            if net is not None:
                net.forward_async(images[start:stop], out[start:stop])   # queue on every device first
        for net in self.nets:
            if net is not None:
                net.sync()
        return out

    def close(self):
        for net in self.nets:
            if net is not None:
                net.close()
