"""Times the detection-decoding kernel alone (device buffers) on two kinds of head tensors:
random-init network outputs (hundreds of candidates pass) and detector-like ones (a few dozen)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dnn-inference-engine_b200"), os.path.join(ROOT, "oracle")]
import _lib  # noqa: E402
import postproc as pp  # noqa: E402


def main():
    B = 512
    for name, pred in [("detector-like", pp.synth_predictions(B, 3)),
                       ("random-head", (np.random.default_rng(0).standard_normal((B, 13, 13, 125)) * 4.5).astype(np.float32))]:
        t0 = time.perf_counter()
        dets, counts = _lib.postprocess(pred, max_det=64)
        t1 = time.perf_counter()
        dets, counts = _lib.postprocess(pred, max_det=64)
        t2 = time.perf_counter()
        print("%-14s B=%d  host-inclusive %.2f ms (first %.2f), boxes/image %.1f, max %d" %
              (name, B, (t2 - t1) * 1e3, (t1 - t0) * 1e3, counts.mean(), counts.max()))


if __name__ == "__main__":
    main()
