"""Short profiling target: builds the YOLOv2-tiny network at the bench micro-batch and runs a few
forwards from device-resident inputs.  Usage: python tools/ncu_target.py [precision] [batch] [forwards]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dnn-inference-engine_b200"))
import numpy as np  # noqa: E402
import _lib  # noqa: E402
from synth import POOLS, make_weights  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "tf32"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 512
forwards = int(sys.argv[3]) if len(sys.argv) > 3 else 3
spec = []
for i, w in enumerate(make_weights(0)):
    blk = {"kernel": w["kernel"], "padding": "SAME", "biases": w["biases"]}
    if "gamma" in w:
        blk.update(mean=w["moving_mean"], variance=w["moving_variance"], gamma=w["gamma"], epsilon=1e-5, leaky=True)
    blk["pool"] = None if POOLS[i] is None else (POOLS[i][0], POOLS[i][1], "SAME")
    spec.append(blk)
net = _lib.Net((batch, 416, 416, 3), spec, precision=precision)
x = np.random.default_rng(0).random((batch, 416, 416, 3), dtype=np.float32)
out = net.forward(x)                      # uploads x once (stays in the network's input buffer)
for _ in range(forwards - 1):
    net.forward(x)
print("ok", out.shape, float(np.abs(out).max()))
