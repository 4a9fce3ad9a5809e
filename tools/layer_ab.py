"""A/B of planner switches on one box: per-block times (alone, L2 flushed) and whole-forward throughput at a micro-batch,
for each environment given as 'NAME=value,NAME=value' arguments ('-' = default plan).
  python tools/layer_ab.py 512 tf32 - DNNB_CONV0_V1=1"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dnn-inference-engine_b200"))
import torch  # noqa: E402

import _lib  # noqa: E402
import bench  # noqa: E402
from synth import make_weights  # noqa: E402

batch, precision = int(sys.argv[1]), sys.argv[2]
spec = bench.yolo_spec(make_weights(0))
x = torch.rand((batch, 416, 416, 3), device="cuda")
y = torch.empty((batch, 13, 13, 125), device="cuda")
stream = torch.cuda.Stream()
for rep in range(2):
    for envs in sys.argv[3:]:
        keys = []
        if envs != "-":
            for kv in envs.split(","):
                k, v = kv.split("=")
                os.environ[k] = v
                keys.append(k)
        net = _lib.Net((batch, 416, 416, 3), spec, precision=precision)
        net.forward_device(x.data_ptr(), None, None)
        net.sync()
        ms = net.time_blocks(2, 8)
        with torch.cuda.stream(stream):
            for _ in range(3):
                net.forward_device(x.data_ptr(), y.data_ptr(), stream.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            n = 40
            for _ in range(n):
                net.forward_device(x.data_ptr(), y.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            e1.synchronize()
        fwd = e0.elapsed_time(e1) / n
        print("%-40s blocks(ms) %s  sum %.3f  forward %.3f ms = %.0f img/s" %
              (envs, " ".join("%.3f" % t for t in ms), sum(ms), fwd, batch / fwd * 1e3), flush=True)
        net.close()
        for k in keys:
            os.environ.pop(k, None)
