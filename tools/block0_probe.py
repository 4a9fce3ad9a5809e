"""Runs block 0 alone (micro-batch from argv; further arguments NAME=value set planner switches for a second run) so
that ncu can capture the conv0 kernel:
  ncu --set full --clock-control none --import-source on -k regex:conv0 -c 2 -o gpurun_out/conv0 python tools/block0_probe.py 512"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dnn-inference-engine_b200"))
import torch  # noqa: E402

import _lib  # noqa: E402
import bench  # noqa: E402
from synth import make_weights  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
spec = bench.yolo_spec(make_weights(0))[:2]
x = torch.rand((batch, 416, 416, 3), device="cuda")
for env in [{}] + [dict([a.split("=")]) for a in sys.argv[2:]]:
    os.environ.update(env)
    net = _lib.Net((batch, 416, 416, 3), spec, precision="tf32")
    net.debug_block_limit(1)
    for _ in range(2):
        net.forward_device(x.data_ptr(), None, None)
    net.sync()
    print(env, net.block_info(0))
    net.close()
    for k in env:
        os.environ.pop(k)
