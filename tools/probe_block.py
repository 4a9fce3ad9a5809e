"""Bring-up probe (GPU box): runs one YOLO conv block through the tcgen05 path and through the
CUDA-core fp32 path of the same library on the same input, and prints where they differ
(overall error, error against spatially shifted references, per-row / per-channel-group counts).
Usage: python tools/probe_block.py LAYER BATCH [tf32|bf16] [nopool]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dnn-inference-engine_b200"))
import _lib  # noqa: E402
from synth import CHANNELS, POOLS, make_weights  # noqa: E402

HW = [416, 208, 104, 52, 26, 13, 13, 13, 13]


def spec(layer, w, pool=True):
    s = {"kernel": w["kernel"], "padding": 'SAME', "biases": w["biases"]}
    if "gamma" in w:
        s.update(mean=w["moving_mean"], variance=w["moving_variance"], gamma=w["gamma"], epsilon=1e-5, leaky=True)
    p = POOLS[layer] if pool else None
    s["pool"] = None if p is None else (p[0], p[1], 'SAME')
    return s


def main():
    layer, batch = int(sys.argv[1]), int(sys.argv[2])
    prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
    pool = not (len(sys.argv) > 4 and sys.argv[4] == "nopool")
    w = make_weights(20 + layer)[layer]
    x = np.random.default_rng(40 + layer).standard_normal((batch, HW[layer], HW[layer], CHANNELS[layer])).astype(np.float32)
    ref_net = _lib.Net(x.shape, [spec(layer, w, pool)], precision="fp32")
    ref = ref_net.forward(x)
    net = _lib.Net(x.shape, [spec(layer, w, pool)], precision=prec)
    info = net.block_info(0)
    got = net.forward(x)
    scale = np.abs(ref).max()
    err = np.abs(got - ref).max() / scale
    print("layer %d batch %d %s pool=%s: %s" % (layer, batch, prec, pool, info))
    print("  normalised error %.4g   (ref absmax %.4g, got absmax %.4g, got nonfinite %d)" %
          (err, scale, np.abs(np.nan_to_num(got)).max(), int((~np.isfinite(got)).sum())))
    tol = 3e-3 if prec == "tf32" else 2e-2
    if err < tol:
        print("  PASS")
        return 0
    bad = np.abs(got - ref) > tol * scale
    print("  FAIL: %.2f%% of elements off" % (100.0 * bad.mean()))
    b, h, wd, c = ref.shape
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            sh = np.roll(ref, (dy, dx), axis=(1, 2))
            inner = (slice(None), slice(1, h - 1), slice(1, wd - 1), slice(None))
            e = np.abs(got[inner] - sh[inner]).max() / scale
            print("    vs ref shifted (dy=%+d, dx=%+d): interior error %.4g" % (dy, dx, e))
    print("    bad per image      :", bad.reshape(b, -1).mean(1).round(3).tolist()[:16])
    print("    bad per row h      :", bad.mean(axis=(0, 2, 3)).round(2).tolist()[:32])
    print("    bad per col w      :", bad.mean(axis=(0, 1, 3)).round(2).tolist()[:32])
    print("    bad per channel/8  :", bad.reshape(b, h, wd, -1, min(8, c)).mean(axis=(0, 1, 2, 4)).round(2).tolist()[:32])
    idx = np.argwhere(bad)[:8]
    for i in idx:
        print("    e.g. at", tuple(i), "got %.5g ref %.5g" % (got[tuple(i)], ref[tuple(i)]))
    zero = (got == 0).mean()
    print("    fraction of exact zeros in got: %.3f" % zero)
    return 1


if __name__ == "__main__":
    sys.exit(main())
