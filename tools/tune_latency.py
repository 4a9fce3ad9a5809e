"""Measures, on the device, the N-tile width of every generic-kernel block under the batch-1 latency plan: coordinate
descent over the blocks (two passes), each candidate a freshly planned network timed over whole forwards (PDL chain +
graph replay, as in production), because a block's best width depends on its neighbours -- a grid that fills every SM
keeps its successor's prologue from starting early.  Prints the planner's modelled choice next to the measured one.
  python tools/tune_latency.py tf32 [forwards]"""
import statistics
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dnn-inference-engine_b200"))
import torch  # noqa: E402

import _lib  # noqa: E402
import bench  # noqa: E402
from synth import make_weights  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
forwards = int(sys.argv[2]) if len(sys.argv) > 2 else 80
spec = bench.yolo_spec(make_weights(0))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
x = torch.rand((1, 416, 416, 3), device=dev)
y = torch.empty((1, 13, 13, 125), device=dev)


def measure(widths):
    net = _lib.Net((1, 416, 416, 3), spec, precision=prec, latency=True, latency_widths=widths)
    plan = [(net.block_info(i)["kernel"], net.block_info(i)["bn_tile"], net.block_info(i)["splits"], net.block_info(i)["cluster_splitk"])
            for i in range(net.num_blocks)]
    ts = []
    with torch.cuda.stream(stream):
        for i in range(10 + forwards):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            net.forward_device(x.data_ptr(), y.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            e1.synchronize()
            if i >= 10:
                ts.append(e0.elapsed_time(e1) * 1e3)
    net.close()
    return statistics.median(ts), plan


n_blocks = len(spec)
base, plan0 = measure([0] * n_blocks)
print("modelled plan: %.1f us  %s" % (base, [(p[1], p[2], p[3]) for p in plan0]))
widths = [0] * n_blocks
best = base
tunable = [i for i in range(n_blocks) if plan0[i][0] in ("umma_im2col", "umma_tiled_pool")]
for sweep in range(2):
    for i in tunable:
        for cand in (256, 128, 64, 32):
            if cand == (widths[i] or plan0[i][1]):
                continue
            trial = list(widths)
            trial[i] = cand
            t, plan = measure(trial)
            if plan[i][1] != cand or plan[i][0] == "simt":
                continue                                    # no kernel of that width: the planner fell back
            mark = ""
            if t < best - 0.3:
                best, widths, mark = t, trial, "  <- kept"
            print("  sweep %d block %d width %3d (splits %d cluster %d): %.1f us%s" % (sweep, i, cand, plan[i][2], plan[i][3], t, mark), flush=True)
t, plan = measure(widths)
print("measured best: widths %s  %.1f us (modelled plan %.1f us)  %s" % (widths, t, base, [(p[1], p[2], p[3]) for p in plan]))
