"""Times one YOLO block alone (L2 flushed between repetitions).  Usage: python tools/time_layer.py LAYER BATCH [precision]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dnn-inference-engine_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import _lib  # noqa: E402
from probe_block import HW, spec  # noqa: E402
from synth import CHANNELS, make_weights  # noqa: E402

layer, batch = int(sys.argv[1]), int(sys.argv[2])
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
pool = not (len(sys.argv) > 4 and sys.argv[4] == "nopool")
w = make_weights(0)[layer]
x = np.random.default_rng(1).random((batch, HW[layer], HW[layer], CHANNELS[layer]), dtype=np.float32)
net = _lib.Net(x.shape, [spec(layer, w, pool)], precision=prec)
net.forward(x)
ms = net.time_blocks(3, 10)
print("layer %d batch %d %s %s: %.4f ms  [%s]" % (layer, batch, prec, net.block_info(0)["kernel"], ms[0],
      " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("DNNB_"))))
