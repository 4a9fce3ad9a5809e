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


class OverlappedGather:
    """The final output gather of the one-process-per-GPU driver, kept off the forwards' critical path: the
    local outputs are double-buffered, and the gather of step k runs on its own stream while the forwards of
    step k + 1 fill the other buffer.  Every gather is still part of the step that produced it: `drain`
    makes the compute stream wait for all of them (bench.py calls it before the closing event).

        g = OverlappedGather(local_shape, dist, device)          # dist: torch.distributed, initialised; or None
        buf = g.begin(stream)      # output tensor for this step, free to overwrite once `stream` reaches this point
        ... launch the forwards of the step on `stream`, writing into buf ...
        g.end(stream)              # enqueue the gather of buf to rank `dst`
        g.drain(stream)            # `stream` waits for every gather so far
        g.gathered                 # on rank dst: list of world tensors, rank r's slice in slot r (None elsewhere)
    """

    def __init__(self, local_shape, dist=None, device=None, dst=0, dtype=None):
        import torch
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.dst = dst
        dtype = dtype or torch.float32
        n_buf = 2 if self.world > 1 else 1
        self.bufs = [torch.empty(local_shape, device=device, dtype=dtype) for _ in range(n_buf)]
        self.gathered = ([torch.empty(local_shape, device=device, dtype=dtype) for _ in range(self.world)]
                         if (self.world > 1 and self.rank == dst) else None)
        self.cuda = self.bufs[0].is_cuda
        self.comm = torch.cuda.Stream(device=device) if (self.world > 1 and self.cuda) else None
        self.done = [None] * n_buf       # event: the gather that last read bufs[b] has finished
        self.k = 0
        self.cur = 0

    def begin(self, stream=None):
        self.cur = self.k % len(self.bufs)
        self.k += 1
        if self.done[self.cur] is not None and stream is not None:
            stream.wait_event(self.done[self.cur])
        return self.bufs[self.cur]

    def end(self, stream=None):
        if self.world == 1:
            return
        import torch
        buf = self.bufs[self.cur]
        if self.comm is None:                               # CPU tensors (gloo tests): plain blocking gather
            self.dist.gather(buf, self.gathered, dst=self.dst)
            return
        ready = torch.cuda.Event()
        ready.record(stream)
        self.comm.wait_event(ready)
        with torch.cuda.stream(self.comm):
            self.dist.gather(buf, self.gathered, dst=self.dst)     # the only exchange of the whole path
            self.done[self.cur] = torch.cuda.Event()
            self.done[self.cur].record(self.comm)

    def drain(self, stream=None):
        if self.comm is not None and stream is not None:
            stream.wait_stream(self.comm)

    def last(self):
        """Local outputs of the most recent step."""
        return self.bufs[self.cur]


def sample_positions(n_local):
    """Positions of a rank's slice that the untimed parity sample re-computes: first, middle, last."""
    return sorted({0, n_local // 2, n_local - 1})


def check_shards(local_out, local_refs, positions, dist, gathered=None, dst=0):
    """Multi-GPU correctness evidence (images are independent, proj3/dnn_openblas.c:59-60, so rank r's slice
    must be exactly what a single GPU computes for those images).  Every rank passes its local outputs
    [n_local, ...], and `local_refs` [len(positions), ...]: the same images pushed one at a time through a
    batch-1 network.  Checks, bit for bit, that (1) the local outputs at `positions` equal the refs and (2) on
    rank dst, slot r of `gathered` holds rank r's refs at those positions -- i.e. the slice landed in the right
    slot with the right values.  Returns {"ranks", "bit_identical", "mismatches"} on rank dst (all ranks get
    bit_identical)."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    flat = local_out.reshape((local_out.shape[0] if local_out.dim() == len(local_refs.shape) else -1,) + tuple(local_refs.shape[1:]))
    ok_local = all(bool(torch.equal(flat[p], local_refs[i])) for i, p in enumerate(positions))
    mism = 0 if ok_local else 1
    if world > 1:
        refs_all = [torch.empty_like(local_refs) for _ in range(world)] if rank == dst else None
        dist.gather(local_refs.contiguous(), refs_all, dst=dst)
        flag = torch.tensor([mism], device=local_refs.device, dtype=torch.int32)
        if rank == dst:
            assert gathered is not None and len(gathered) == world
            for r in range(world):
                g = gathered[r].reshape((-1,) + tuple(local_refs.shape[1:]))
                for i, p in enumerate(positions):
                    if not bool(torch.equal(g[p], refs_all[r][i])):
                        flag += 1
        dist.all_reduce(flag)
        mism = int(flag.item())
    return {"ranks": world, "bit_identical": mism == 0, "mismatches": mism, "positions_per_rank": list(positions)}


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
        if not (isinstance(images, np.ndarray) and images.dtype == np.float32 and images.flags["C_CONTIGUOUS"]
                and images.shape[0] == self.total):
            raise ValueError("images must be a C-contiguous float32 array of %d images" % self.total)
        if out is None:
            out = np.empty(self.out_shape, dtype=np.float32)
        if not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
                and tuple(out.shape) == tuple(self.out_shape)):
            raise ValueError("out must be a C-contiguous float32 array of shape %s" % (self.out_shape,))
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
