"""Batch-1 latency plan, block by block: the plan of every block and the in-situ share of each (median latency of
forwards cut after k blocks, PDL chain + graph replay as in production).  Extra arguments NAME=value set planner
switches for further runs.   python tools/batch1_blocks.py tf32 [DNNB_NO_CLUSTER_SPLITK=1 ...]"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dnn-inference-engine_b200"))
import torch  # noqa: E402

import _lib  # noqa: E402
import bench  # noqa: E402
from synth import make_weights  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
spec = bench.yolo_spec(make_weights(0))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
x = torch.rand((1, 416, 416, 3), device=dev)
y = torch.empty((1, 13, 13, 125), device=dev)


def times(net, n=60, warm=8):
    ts = []
    with torch.cuda.stream(stream):
        for i in range(warm + n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            net.forward_device(x.data_ptr(), y.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            e1.synchronize()
            if i >= warm:
                ts.append(e0.elapsed_time(e1) * 1e3)
    return ts


for env in [{}] + [dict([a.split("=")]) for a in sys.argv[2:]]:
    os.environ.update(env)
    net = _lib.Net((1, 416, 416, 3), spec, precision=prec, latency=True)
    ts = times(net)
    print(env or "default", "median %.1f us  min %.1f us  launches %d" % (statistics.median(ts), min(ts), net.launches_per_forward()))
    prev = 0.0
    for k in range(1, net.num_blocks + 1):
        net.debug_block_limit(k)
        med = statistics.median(times(net, 30, 5))
        bi = net.block_info(k - 1)
        print("  block %d: +%5.1f us  %-16s grid %3d splits %2d cluster %d pool_sep %d launches %d" %
              (k - 1, med - prev, bi["kernel"], bi["grid"], bi["splits"], bi["cluster_splitk"], bi["pool_separate"], bi["launches"]))
        prev = med
    net.debug_block_limit(0)
    net.close()
    for k in env:
        os.environ.pop(k)
