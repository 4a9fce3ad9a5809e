"""The measured latency plan from the command line: `_lib.tune_latency_widths` (coordinate descent over the blocks'
N-tile widths, every candidate a freshly planned network timed over whole graph-replayed forwards -- what
`DnnInferenceEngine(latency="tune")` runs) with every candidate printed, next to the planner's modelled choice.
  python tools/tune_latency.py tf32 [forwards per candidate]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dnn-inference-engine_b200"))

import _lib  # noqa: E402
import bench  # noqa: E402
from synth import make_weights  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
forwards = int(sys.argv[2]) if len(sys.argv) > 2 else 60
spec = bench.yolo_spec(make_weights(0))
shape = (1, 416, 416, 3)


def plan_of(widths):
    net = _lib.Net(shape, spec, precision=prec, latency=True, latency_widths=widths)
    info = [net.block_info(i) for i in range(net.num_blocks)]
    us = net.time_forward(10, forwards)
    net.close()
    return us, [(b["bn_tile"], b["splits"], int(b["cluster_splitk"])) for b in info]


us0, plan0 = plan_of(None)
print("modelled plan  %.1f us  (N-tile width, K slices, cluster) per block: %s" % (us0, plan0), flush=True)
res = _lib.tune_latency_widths(shape, spec, prec, iters=forwards, log=lambda s: print("  " + s, flush=True))
us1, plan1 = plan_of(res["widths"])
print("measured plan  %.1f us  widths %s  %s" % (us1, res["widths"], plan1))
