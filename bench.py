#!/usr/bin/env python
"""bench.py -- YOLOv2-tiny 416x416 images/s on 1/2/4/8 B200 (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            (N=1)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3]): a global batch of 4096 synthetic images per step, cut into
contiguous slices of 4096/N per GPU (weights replicated, no collective on the hot path, one
final output gather to rank 0), each slice run as micro-batches of 512 through the resident
fused network.  One "step" = one pass of the hot path over the 4096 images.

  value  : images/s with the inputs already resident in HBM (device-timed, max over ranks)
  e2e    : images/s through the host-buffer C-ABI call (pinned host in -> H2D -> 9 fused blocks ->
           D2H -> pinned host out), copies inside the timed region
  roofline: the dominant kernel (block 7, 3x3x1024->1024 on 13x13, 42% of the summed per-layer
           roofline), timed in situ with CUDA events around its launches inside the timed steps
  cpu_baseline: the reference's own OpenBLAS engine (oracle/_ref, compiled unmodified) timed on
           this box's host cores on a bounded sample -- a reported baseline, not the target; its
           AVX2+pthread engine alongside (cpu_baseline.avx_pthread)
  batch1_latency: BASELINE configs[1], the device-resident batch-1 forward under the latency plan

  --impl reference : times that CPU reference alone (rank 0), same metric/unit/config.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "dnn-inference-engine_b200")
for p in (PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
# ncu --set full capture (profiles/), scaled to the bench micro-batch; None where not captured
NCU_TRAFFIC = {"tf32": 1286576000, "bf16": None}   # profiles/r01_ncu_block7_full_raw.csv (main launch), micro-batch 512
METRIC = "YOLOv2-tiny 416x416 images/sec"
UNIT = "images/s"
GFLOP_PER_IMAGE = [0.1495, 0.3987, 0.3987, 0.3987, 0.3987, 0.3987, 1.5949, 3.1898, 0.0433]
HW_IN = [416, 208, 104, 52, 26, 13, 13, 13, 13]
HW_OUT = [208, 104, 52, 26, 13, 13, 13, 13, 13]
CH = [3, 16, 32, 64, 128, 256, 512, 1024, 1024, 125]
KS = [3, 3, 3, 3, 3, 3, 3, 3, 1]


def algorithmic_bytes(layer, batch, act_bytes):
    """unpadded input once + post-pool output once + weights once per batch + scale/shift.
    The network input is always fp32 and so is the final output."""
    e_in = 4 if layer == 0 else act_bytes
    e_out = 4 if layer == 8 else act_bytes
    return (batch * (HW_IN[layer] ** 2 * CH[layer] * e_in + HW_OUT[layer] ** 2 * CH[layer + 1] * e_out)
            + KS[layer] ** 2 * CH[layer] * CH[layer + 1] * act_bytes + 2 * CH[layer + 1] * 4)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled every
    few ms from a thread (nvidia-smi's loop is too coarse for a 60 ms region); nvidia-smi as fallback."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period_s=0.004):
        self.index, self.period, self.samples, self.stop_flag, self.thread, self.err = index, period_s, [], False, None, None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, mx, rs))
                time.sleep(self.period)
        except Exception as e:   # noqa: BLE001
            self.err = repr(e)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.samples:
            return self._smi_once()
        sm = [s[0] for s in self.samples]
        mask = 0
        for s in self.samples:
            mask |= s[2]
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": self.samples[0][1], "samples": len(sm),
                "reasons": sorted(k for k, bit in self.REASONS.items() if mask & bit), "source": "nvml"}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b = [float(v) for v in out.strip().split(",")]
            return {"sm_mhz": a, "sm_max_mhz": b, "samples": 1, "reasons": [], "source": "nvidia-smi after the region (nvml failed: %s)" % self.err}
        except Exception:   # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock sampling unavailable: %s" % self.err]}


def openblas_threads():
    try:
        import build as oracle_build
        lib = ctypes.CDLL(oracle_build.scipy_openblas())
        return int(lib.scipy_openblas_get_num_threads())
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_rate(n_images, weights, seed=100, backend=None):
    """Times the reference CPU engine on n_images (batch 1 each, fresh graph per image, wall clock
    around the run only, as proj3/__init__.py:24-26 does).  Returns (images/s, kind, cores, seconds).
    backend: "ref_openblas" (default when built), "ref_avx" (proj3/dnn_avx.c, 4 pthreads), "oracle"."""
    import chain
    from synth import build_yolov2tiny, make_images
    if backend is None:
        backend = "ref_openblas" if chain.available("ref_openblas") else "oracle"
    kind = "port" if backend == "oracle" else "reference"
    if backend == "oracle":
        import build as oracle_build
        oracle_build.build_oracle()
    total = 0.0
    for i in range(n_images):
        img = make_images(1, seed=seed, start=i)
        c = chain.Chain(backend)
        build_yolov2tiny(c, img.shape, weights)
        t0 = time.perf_counter()
        c.run(img)
        total += time.perf_counter() - t0
    cores = openblas_threads() if backend == "ref_openblas" else 4 if backend == "ref_avx" else 1
    return n_images / total, kind, cores, total


def batch1_latency(_lib, spec, precision, device, forwards=100, warmup=10):
    """BASELINE configs[1]: batch-1 latency of the device-resident forward under the latency plan
    (DNNB_LATENCY_MODE=1: split-K over co-resident CTAs, kernels chained by programmatic dependent launch,
    the launch sequence replayed from a CUDA graph).  Each forward is timed on its own with CUDA events."""
    import torch
    net = _lib.Net((1, 416, 416, 3), spec, precision=precision, device=device, latency=True)
    dev = torch.device("cuda", device)
    stream = torch.cuda.current_stream(dev)
    x = torch.rand((1, 416, 416, 3), device=dev)
    y = torch.empty((1, 13, 13, 125), device=dev)
    ts = []
    for i in range(warmup + forwards):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        net.forward_device(x.data_ptr(), y.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        e1.synchronize()
        if i >= warmup:
            ts.append(e0.elapsed_time(e1) * 1e3)
    launches = net.launches_per_forward()
    # the same from pinned host buffers through the host-facing calls (H2D + forward + D2H + sync, wall clock)
    hin = _lib.PinnedArray((1, 416, 416, 3))
    hu8 = _lib.PinnedArray((1, 416, 416, 3), dtype=np.uint8)
    hout = _lib.PinnedArray((1, 13, 13, 125))
    hin.array[...] = x.cpu().numpy()
    hu8.array[...] = np.random.default_rng(7).integers(0, 256, size=hu8.array.shape, dtype=np.uint8)
    host = {}
    for name, call in (("float32", lambda: net.forward(hin.array, hout.array)),
                       ("uint8_frame", lambda: net.forward_u8(hu8.array, True, hout.array))):
        tw = []
        for i in range(warmup + forwards):
            t0 = time.perf_counter()
            call()
            if i >= warmup:
                tw.append((time.perf_counter() - t0) * 1e6)
        host[name] = round(statistics.median(tw), 1)
    hin.free(); hu8.free(); hout.free()
    net.close()
    return {"median_us": round(statistics.median(ts), 1), "min_us": round(min(ts), 1), "forwards": forwards,
            "launches_per_forward": launches, "precision": precision,
            "e2e_host_median_us": host,   # pinned host image in -> head tensor back in host memory
            "plan": "DNNB_LATENCY_MODE=1: split-K, programmatic dependent launch chain, CUDA graph replay; input resident in HBM"}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    from synth import make_weights
    weights = make_weights(0)
    sample = args.ref_sample
    for _ in range(args.warmup):
        cpu_reference_rate(sample, weights)
    t_total, n_total, kind, cores = 0.0, 0, "port", 1
    for s in range(args.steps):
        rate, kind, cores, secs = cpu_reference_rate(sample, weights, seed=100 + s * sample)
        t_total += secs
        n_total += sample
    value = n_total / t_total
    desc = "%d images per step, batch 1 each, proj3 OpenBLAS engine compiled unmodified (oracle/_ref)" % sample \
        if kind == "reference" else "%d images per step through oracle/oracle.c" % sample
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cpu=True),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    args.out.emit(json.dumps(line))
    return 0


def workload_config(args, cpu=False):
    return {"workload": "YOLOv2-tiny 416x416x3, random-init weights, global batch %d sharded over %d GPU(s), "
                        "micro-batch %d (BASELINE configs[3])" % (args.global_batch, args.gpus, args.micro_batch),
            "global_batch": args.global_batch, "micro_batch": args.micro_batch,
            "precision": "f32 (CPU)" if cpu else args.precision,
            "l2": "inputs larger than L2 (%.1f GB read per step per GPU)" %
                  (args.global_batch / max(1, args.gpus) * 416 * 416 * 3 * 4 / 1e9),
            "parallelism": "batch-sharded dp%d, weights replicated, final output gather (on its own stream, "
                           "overlapping the next step's forwards)" % args.gpus}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import _lib
    from synth import CHANNELS, KSIZES, POOLS, make_weights

    rank, local_rank, world = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:   # plain `python bench.py --gpus N`: relaunch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29591", os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd, stdout=args.out.saved)
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    assert args.global_batch % world == 0
    n_local = args.global_batch // world
    mb = min(args.micro_batch, n_local)
    assert n_local % mb == 0
    n_mb = n_local // mb

    weights = make_weights(0)
    spec = []
    for i, w in enumerate(weights):
        blk = {"kernel": w["kernel"], "padding": "SAME", "biases": w["biases"]}
        if "gamma" in w:
            blk.update(mean=w["moving_mean"], variance=w["moving_variance"], gamma=w["gamma"], epsilon=1e-5, leaky=True)
        blk["pool"] = None if POOLS[i] is None else (POOLS[i][0], POOLS[i][1], "SAME")
        spec.append(blk)
    net = _lib.Net((mb, 416, 416, 3), spec, precision=args.precision, device=local_rank)
    plan = [net.block_info(i) for i in range(net.num_blocks)]
    act_bytes = 2 if args.precision == "bf16" else 4

    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    x = torch.rand((n_mb, mb, 416, 416, 3), device=dev, dtype=torch.float32, generator=gen)
    # N > 1: outputs are double-buffered so that the gather of step k (on its own stream) overlaps the
    # forwards of step k + 1; every gather is inside the timed region (the closing fence waits for it)
    n_buf = 2 if world > 1 else 1
    ys = [torch.empty((n_mb, mb, 13, 13, 125), device=dev, dtype=torch.float32) for _ in range(n_buf)]
    y = ys[0]
    gathered = [torch.empty_like(y) for _ in range(world)] if (world > 1 and rank == 0) else None
    # a real (non-default) stream: the C ABI treats a NULL stream as "the network's own stream"
    stream = torch.cuda.Stream(device=dev)
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    gather_done = [None] * n_buf          # event: the gather that last read ys[b] has finished
    state = {"k": 0}
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)

    def step():
        b = state["k"] % n_buf
        state["k"] += 1
        if gather_done[b] is not None:
            stream.wait_event(gather_done[b])            # ys[b] is free again
        for j in range(n_mb):
            net.forward_device(x[j].data_ptr(), ys[b][j].data_ptr(), stream.cuda_stream)
        if world > 1:
            ready = torch.cuda.Event()
            ready.record(stream)
            comm.wait_event(ready)
            with torch.cuda.stream(comm):
                dist.gather(ys[b], gathered, dst=0)      # the only exchange: final output gather
                gather_done[b] = torch.cuda.Event()
                gather_done[b].record(comm)

    def fence():
        if comm is not None:
            stream.wait_stream(comm)                     # events recorded on `stream` after this see every gather
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # burst tensor peak first, on a cool part (best of 3 calls x 5 launches): denominator for kernels timed alone
    tensor_peak = max(_lib.measure_mma_peak(args.precision, local_rank) for _ in range(3)) if rank == 0 else None

    for _ in range(max(3, args.warmup)):
        step()
    fence()
    dom = 7
    net.profile_block(dom)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    if comm is not None:
        stream.wait_stream(comm)                         # the last gather belongs to the timed region
    e1.record(stream)
    fence()
    y = ys[(state["k"] - 1) % n_buf]                     # outputs of the last step
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    dom_ms, dom_n = net.profile_read()
    net.profile_block(-1)
    assert bool(torch.isfinite(y).all()), "non-finite outputs"
    value = args.global_batch * args.steps / (total_ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, copies timed)
    e2e = None
    if not args.no_e2e:
        hin = _lib.PinnedArray((n_mb, mb, 416, 416, 3))
        hout = _lib.PinnedArray((n_mb, mb, 13, 13, 125))
        for j in range(n_mb):
            hin.array[j] = x[j].cpu().numpy()

        def e2e_step():
            for j in range(n_mb):
                net.forward_async(hin.array[j], hout.array[j])
            net.sync()

        e2e_step()
        k_e2e = max(2, min(args.steps, 5))
        fence()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        fence()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ref_out = y[n_mb - 1].cpu().numpy()
        same = bool(np.array_equal(ref_out, hout.array[n_mb - 1]))
        e2e = {"value": args.global_batch * k_e2e / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(args.global_batch * 416 * 416 * 3 * 4),
               "d2h_bytes_per_step": int(args.global_batch * 13 * 13 * 125 * 4),
               "steps": k_e2e, "matches_device_path": same,
               "note": "per micro-batch: cudaMemcpyAsync H2D -> 9 fused blocks -> cudaMemcpyAsync D2H, on three streams so the "
                       "copies of neighbouring micro-batches overlap the kernels (double-buffered input and output)"}
        hin.free()
        # the same through the uint8-frame entry point: 3 bytes per pixel over PCIe, /255 on the device
        hu8 = _lib.PinnedArray((n_mb, mb, 416, 416, 3), dtype=np.uint8)
        u8 = hu8.array
        u8[...] = np.random.default_rng(5 + rank).integers(0, 256, size=(mb, 416, 416, 3), dtype=np.uint8)[None]

        def e2e_u8_step():
            for j in range(n_mb):
                net.forward_u8_async(u8[j], hout.array[j], True)
            net.sync()

        e2e_u8_step()
        fence()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_u8_step()
        fence()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e["uint8_frames"] = {"value": args.global_batch * k_e2e / float(dt.item()), "unit": UNIT,
                               "h2d_bytes_per_step": int(args.global_batch * 416 * 416 * 3),
                               "note": "dnnb_net_forward_u8_async: uint8 BGR frames in, float32(u8/255.) + BGR->RGB "
                                       "(proj3/__init__.py:8-12 minus the resize) as one extra kernel on the device"}
        # frames in, label boxes out (SURVEY 8f1 + 8f3): the head tensor never leaves the device
        max_det = 32
        net.set_detect(max_det=max_det)
        hdets = _lib.PinnedArray((n_mb, mb, max_det), dtype=_lib.DET_DTYPE)
        hcnt = _lib.PinnedArray((n_mb, mb), dtype=np.int32)

        def e2e_boxes_step():
            for j in range(n_mb):
                net.detect_async(u8[j], hdets.array[j], hcnt.array[j], True)
            net.sync()

        e2e_boxes_step()
        fence()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_boxes_step()
        fence()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e["uint8_frames_to_boxes"] = {
            "value": args.global_batch * k_e2e / float(dt.item()), "unit": UNIT,
            "h2d_bytes_per_step": int(args.global_batch * 416 * 416 * 3),
            "d2h_bytes_per_step": int(args.global_batch * (max_det * _lib.DET_DTYPE.itemsize + 4)),
            "boxes_per_image": float(np.minimum(hcnt.array, max_det).mean()),
            "note": "dnnb_net_detect_u8_async: as uint8_frames, plus the reference's postprocessing (decode, 0.3 "
                    "threshold, NMS; proj3/yolov2tiny.py:94-226) as one more kernel; only boxes cross PCIe. "
                    "Random-init weights make far more candidates pass than a trained head would"}
        hu8.free(); hout.free(); hdets.free(); hcnt.free()

    # ---- per-layer roofline (micro-batch, L2 flushed between repetitions)
    peaks = measured_peaks()
    layers = None
    if rank == 0:
        layer_ms = net.time_blocks(2, 5)                                                   # before the 3 s burn below
        tensor_peak_sustained = _lib.measure_mma_peak(args.precision, local_rank, 3.0)   # under the power cap
        layers = []
        for i, t in enumerate(layer_ms):
            fl = GFLOP_PER_IMAGE[i] * 1e9 * mb
            by = algorithmic_bytes(i, mb, act_bytes)
            t_mem, t_tc = by / (peaks["hbm_gbs"] * 1e9), fl / (tensor_peak * 1e12)
            bound = "hbm" if t_mem >= t_tc else "tensor"
            layers.append({"layer": i, "kernel": plan[i]["kernel"], "ms": round(t, 4), "bound": bound,
                           "gbs": round(by / (t * 1e-3) / 1e9, 1), "tflops": round(fl / (t * 1e-3) / 1e12, 1),
                           "frac": round(max(t_mem, t_tc) / (t * 1e-3), 4)})

    if rank == 0:
        dom_flops = GFLOP_PER_IMAGE[dom] * 1e9 * mb
        achieved = dom_flops / (dom_ms / max(dom_n, 1) * 1e-3) / 1e12 if dom_n else None
        roofline = {"bound": "tensor", "kernel": "block 7 (13x13x1024 -> 1024, 3x3): conv_pair_kernel (tcgen05.mma.cta_group::2, im2col TMA) on whole "
                              "waves + a conv_umma_kernel tail launch over the last rows; both are inside the timed interval",
                    "achieved": achieved, "peak": tensor_peak_sustained, "unit": "TFLOP/s",
                    "frac": achieved / tensor_peak_sustained if achieved else None, "traffic": NCU_TRAFFIC.get(args.precision),
                    "peak_burst": tensor_peak, "frac_of_burst": achieved / tensor_peak if achieved else None,
                    "peak_source": "measured on this box: tcgen05 kind::%s issue loop (M128 N256, operands in smem), sustained "
                                   "for 3 s (the kernel is timed inside a long step); burst figure alongside. "
                                   "MEASURED_PEAKS.json (%s) holds only bf16 cuBLAS: %.1f burst / %.1f sustained TFLOP/s" %
                                   ("tf32" if args.precision == "tf32" else "f16/bf16", peaks["source"],
                                    peaks["bf16_tflops"], peaks["bf16_tflops_sustained"]),
                    "launches_timed": dom_n, "avg_launch_ms": dom_ms / max(dom_n, 1),
                    "share_of_step": dom_ms / total_ms if total_ms else None}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, kind, cores, secs = cpu_reference_rate(args.cpu_sample, weights)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": "%d images, batch 1 each, %.1f s of CPU time; proj3 OpenBLAS engine compiled unmodified, "
                             "host has %d cpus" % (args.cpu_sample, secs, os.cpu_count() or 0)}
            import chain
            if chain.available("ref_avx"):      # the reference's second CPU engine (proj3/dnn_avx.c, P_THREADS 4)
                rate2, _, cores2, secs2 = cpu_reference_rate(args.cpu_sample, weights, backend="ref_avx")
                cpu["avx_pthread"] = {"value": rate2, "unit": UNIT, "cores": cores2, "kind": "reference",
                                      "sample": "%d images, batch 1 each, %.1f s; proj3 AVX2+pthread engine compiled "
                                                "unmodified" % (args.cpu_sample, secs2)}
        b1 = batch1_latency(_lib, spec, args.precision, local_rank) if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic", "config": workload_config(args),
            "e2e": e2e, "gpu_launches": net.launches_per_forward() * n_mb * args.steps,   # timed region of `value`
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "batch1_latency": b1, "layers": layers,
            "sum_layer_roofline_frac": (sum(max(algorithmic_bytes(i, mb, act_bytes) / (peaks["hbm_gbs"] * 1e9),
                                                 GFLOP_PER_IMAGE[i] * 1e9 * mb / (tensor_peak * 1e12)) for i in range(9))
                                        / (sum(l["ms"] for l in layers) * 1e-3)) if layers else None,
        }
        args.out.emit(json.dumps(line))
    net.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


class StdoutGuard:
    """Everything but the final JSON line goes to stderr: libraries (NCCL's version banner, for one)
    write to fd 1 behind Python's back, and the contract is ONE line on stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="tf32", choices=["tf32", "bf16"])
    ap.add_argument("--global-batch", type=int, default=4096)
    ap.add_argument("--micro-batch", type=int, default=512)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=16, help="images timed for cpu_baseline (about 10-30 s)")
    ap.add_argument("--ref-sample", type=int, default=2, help="images per step of --impl reference")
    args = ap.parse_args()
    args.out = StdoutGuard()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
