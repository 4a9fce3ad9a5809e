"""Where the micro-seconds of the batch-1 chain go: per-CTA %globaltimer stamps of the generic tcgen05 kernel
(blocks 3..8 under the latency plan), on one common time axis (ns from the earliest stamp of the forward).
  python tools/batch1_timeline.py tf32 [NAME=value ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dnn-inference-engine_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import _lib  # noqa: E402
import bench  # noqa: E402
from synth import make_weights  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
for a in sys.argv[2:]:
    k, v = a.split("=")
    os.environ[k] = v
spec = bench.yolo_spec(make_weights(0))
x = torch.rand((1, 416, 416, 3), device="cuda")
y = torch.empty((1, 13, 13, 125), device="cuda")
net = _lib.Net((1, 416, 416, 3), spec, precision=prec, latency=True)
for _ in range(5):
    net.forward_device(x.data_ptr(), y.data_ptr(), None)
net.sync()
blocks = [i for i in range(net.num_blocks) if net.block_info(i)["kernel"] == "umma_im2col"]
for i in blocks:
    net.debug_timeline(i)
for _ in range(3):                                  # the third forward is a graph replay again
    net.forward_device(x.data_ptr(), y.data_ptr(), None)
net.sync()
stamps = {i: net.debug_timeline(i, read=True).astype(np.int64) for i in blocks}
t0 = min(int(s[:, 0][s[:, 0] > 0].min()) for s in stamps.values())
names = ["entry", "prologue", "pred done", "1st operands", "last MMA", "acc done", "siblings seen", "stored", "partial written", "partials read"]
order = [0, 1, 2, 3, 4, 5, 8, 6, 9, 7]
for i in blocks:
    s = stamps[i] - t0
    bi = net.block_info(i)
    print("block %d  grid %d splits %d cluster %d" % (i, bi["grid"], bi["splits"], bi["cluster_splitk"]))
    for j in order:
        nm = names[j]
        col = s[:, j][stamps[i][:, j] > 0]
        if col.size:
            print("   %-14s min %7.2f  median %7.2f  max %7.2f us" % (nm, col.min() / 1e3, np.median(col) / 1e3, col.max() / 1e3))
net.close()
