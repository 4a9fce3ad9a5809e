"""Measures the BASELINE.json configs that bench.py's headline line does not cover:
  configs[1]  batch-1 latency (median of N CUDA-event-timed forwards, input resident) vs the CPU engines
  configs[2]  batch-256 throughput, TF32 and BF16
  configs[4]  per-layer microbench at batch 256 and batch 1 (GB/s, TFLOP/s, fraction of roofline)
Writes one JSON document to stdout.  Usage: python tools/configs_bench.py [--forwards 200]"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dnn-inference-engine_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import _lib  # noqa: E402
import bench  # noqa: E402
from synth import POOLS, build_yolov2tiny, make_images, make_weights  # noqa: E402


def make_net(batch, precision):
    spec = []
    for i, w in enumerate(make_weights(0)):
        blk = {"kernel": w["kernel"], "padding": "SAME", "biases": w["biases"]}
        if "gamma" in w:
            blk.update(mean=w["moving_mean"], variance=w["moving_variance"], gamma=w["gamma"], epsilon=1e-5, leaky=True)
        blk["pool"] = None if POOLS[i] is None else (POOLS[i][0], POOLS[i][1], "SAME")
        spec.append(blk)
    return _lib.Net((batch, 416, 416, 3), spec, precision=precision)


def device_forward_times(net, batch, forwards, warmup=10):
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    x = torch.rand((batch, 416, 416, 3), device=dev)
    y = torch.empty((batch, 13, 13, 125), device=dev)
    torch.cuda.synchronize()
    times = []
    with torch.cuda.stream(stream):
        for i in range(warmup + forwards):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            net.forward_device(x.data_ptr(), y.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            e1.synchronize()
            if i >= warmup:
                times.append(e0.elapsed_time(e1))
    return times


def layer_table(net, batch, precision, peaks, tensor_peak):
    act = 2 if precision == "bf16" else 4
    rows = []
    for i, t in enumerate(net.time_blocks(3, 10)):
        fl = bench.GFLOP_PER_IMAGE[i] * 1e9 * batch
        by = bench.algorithmic_bytes(i, batch, act)
        t_mem, t_tc = by / (peaks["hbm_gbs"] * 1e9), fl / (tensor_peak * 1e12)
        rows.append({"layer": i, "kernel": net.block_info(i)["kernel"], "us": round(t * 1e3, 2),
                     "bound": "hbm" if t_mem >= t_tc else "tensor", "gbs": round(by / (t * 1e-3) / 1e9, 1),
                     "tflops": round(fl / (t * 1e-3) / 1e12, 2), "roofline_us": round(max(t_mem, t_tc) * 1e6, 2),
                     "frac": round(max(t_mem, t_tc) / (t * 1e-3), 4)})
    return rows


def cpu_engines(n_runs):
    import chain
    weights = make_weights(0)
    out = {}
    for backend in ("ref_openblas", "ref_avx"):
        if not chain.available(backend):
            continue
        ts = []
        for i in range(n_runs):
            img = make_images(1, seed=100 + i)
            c = chain.Chain(backend)
            build_yolov2tiny(c, img.shape, weights)
            t0 = time.perf_counter()
            c.run(img)
            ts.append(time.perf_counter() - t0)
        out[backend] = {"median_s": statistics.median(ts), "runs": n_runs,
                        "threads": bench.openblas_threads() if backend == "ref_openblas" else 4}
    out["host_cpus"] = os.cpu_count()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--forwards", type=int, default=200)
    ap.add_argument("--cpu-runs", type=int, default=5)
    args = ap.parse_args()
    peaks = bench.measured_peaks()
    doc = {"peaks": peaks, "tensor_peak_tflops": {}}
    for prec in ("tf32", "bf16"):
        doc["tensor_peak_tflops"][prec] = max(_lib.measure_mma_peak(prec) for _ in range(3))
    # configs[1]: batch-1 latency, default (deterministic) plan and DNNB_LATENCY_MODE=1 (split-K with atomics)
    doc["batch1_latency"] = {}
    for mode in ("default", "latency_mode"):
        if mode == "latency_mode":
            os.environ["DNNB_LATENCY_MODE"] = "1"
        for prec in ("tf32", "bf16"):
            net = make_net(1, prec)
            ts = device_forward_times(net, 1, args.forwards)
            doc["batch1_latency"][prec + "_" + mode] = {
                "median_us": round(statistics.median(ts) * 1e3, 2), "min_us": round(min(ts) * 1e3, 2),
                "p90_us": round(sorted(ts)[int(0.9 * len(ts))] * 1e3, 2), "forwards": len(ts),
                "layers": layer_table(net, 1, prec, peaks, doc["tensor_peak_tflops"][prec])}
            net.close()
    os.environ.pop("DNNB_LATENCY_MODE", None)
    # configs[2] and the per-layer table at batch 256
    doc["batch256"] = {}
    for prec in ("tf32", "bf16"):
        net = make_net(256, prec)
        ts = device_forward_times(net, 256, 30, warmup=5)
        med = statistics.median(ts)
        doc["batch256"][prec] = {"images_per_s": round(256 / (med * 1e-3)), "ms_per_forward": round(med, 4),
                                 "layers": layer_table(net, 256, prec, peaks, doc["tensor_peak_tflops"][prec])}
        net.close()
    doc["cpu_engines_batch1"] = cpu_engines(args.cpu_runs)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
