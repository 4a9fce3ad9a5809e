"""Pair-kernel (cta_group::2) bring-up check: the network with and without CTA pairs must agree bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dnn-inference-engine_b200"), os.path.join(ROOT, "oracle")]
import dnn_b200  # noqa: E402
from synth import build_yolov2tiny, make_images, make_weights  # noqa: E402


def run(batch, precision, env):
    for k in ("DNNB_NO_PAIR", "DNNB_PAIR_MT", "DNNB_PAIR_MIN_TILES"):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = dnn_b200.DnnGraphBuilder()
    build_yolov2tiny(g, [batch, 416, 416, 3], make_weights(0))
    eng = dnn_b200.DnnInferenceEngine(g, False, precision=precision)
    out = eng.run(make_images(batch, seed=100)).copy()
    plan = [(p["kernel"], p["grid"], p["launches"]) for p in eng.kernel_plan()]
    outs = [eng._net.block_output(i) for i in (5, 6, 7)]
    eng.close()
    return out, plan, outs


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    for precision in ("tf32", "bf16"):
        ref, plan0, blocks0 = run(batch, precision, {"DNNB_NO_PAIR": "1"})
        for env in ({}, {"DNNB_PAIR_MT": "1"}, {"DNNB_PAIR_MT": "2"}):
            env = dict(env, DNNB_PAIR_MIN_TILES="1")
            out, plan, blocks = run(batch, precision, env)
            same = [bool(np.array_equal(a, b)) for a, b in zip(blocks0, blocks)]
            print(precision, env, "blocks 5-7 identical:", same, "final identical:", bool(np.array_equal(out, ref)),
                  "max diff", float(np.abs(out - ref).max()), plan[5:8], flush=True)


if __name__ == "__main__":
    main()
