"""GPU parity tests for the on-device input preprocessing (`resize_input`,
proj3/__init__.py:8-12 = cv2.resize + /255. + BGR->RGB + float32) against oracle/resize.py and the
cv2 outputs stored in tests/golden/resize.npz.  Integer work: everything is bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
import _lib
import dnn_b200
import resize as rs
from synth import build_yolov2tiny, make_weights

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(GOLDEN, "resize.npz"))


def as_float_rgb(u8_bgr):
    return np.asarray((u8_bgr / 255.)[..., ::-1], dtype=np.float32)


def test_op_level_resize_input_against_cv2_goldens():
    for i, (h, w, d) in enumerate(G["cases"]):
        got = _lib.resize_input(G["src_%d" % i], hw=int(d), bgr=True)
        np.testing.assert_array_equal(got, as_float_rgb(G["dst_%d" % i]), err_msg="case %d" % i)
        got_rgb = _lib.resize_input(G["src_%d" % i], hw=int(d), bgr=False)
        np.testing.assert_array_equal(got_rgb, np.asarray(G["dst_%d" % i] / 255., dtype=np.float32))
    big = _lib.resize_input(G["big_src"], hw=416)
    u8 = np.rint(big[..., ::-1] * 255.0).astype(np.int64)
    assert int(u8.sum()) == int(G["big_sum"])
    np.testing.assert_array_equal(u8[::13, ::7], G["big_sample"])


def test_op_level_resize_input_batched_random_sizes_against_oracle():
    rng = np.random.default_rng(8)
    for h, w in [(480, 640), (375, 500), (208, 208), (1080, 1920), (417, 415), (2, 3)]:
        frames = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
        got = _lib.resize_input(frames, hw=416)
        for n in range(3):
            np.testing.assert_array_equal(got[n], rs.resize_input(frames[n], 416), err_msg="%dx%d image %d" % (h, w, n))
    try:
        import cv2
    except ImportError:
        return
    frame = rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    np.testing.assert_array_equal(_lib.resize_input(frame), as_float_rgb(cv2.resize(frame, (416, 416))))


def test_network_on_frames_equals_network_on_the_reference_preprocessing():
    B = 2
    g = dnn_b200.DnnGraphBuilder()
    build_yolov2tiny(g, [B, 416, 416, 3], make_weights(0))
    eng = dnn_b200.DnnInferenceEngine(g, False, precision="tf32")
    eng.run(np.zeros((B, 416, 416, 3), np.float32))       # the resident network is built on the first run
    net = eng._net
    rng = np.random.default_rng(12)
    for h, w in [(480, 640), (300, 200), (416, 416), (600, 800)]:        # grows the staging buffers, swaps tables
        frames = rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)
        x = np.stack([rs.resize_input(f, 416) for f in frames])
        want = eng.run(x).copy()
        got = net.forward_frames(frames, bgr=True)
        np.testing.assert_array_equal(got, want, err_msg="%dx%d" % (h, w))
    net.set_detect(max_det=_lib.CANDIDATES)
    dets, counts = net.detect(frames, bgr=True)                           # 600x800 frames -> boxes
    dets2, counts2 = _lib.postprocess(want, max_det=_lib.CANDIDATES)
    np.testing.assert_array_equal(counts, counts2)
    np.testing.assert_array_equal(dets, dets2)
    eng.close()
