"""CPU tests: oracle/resize.py (restatement of cv2.resize's 8-bit bilinear path, which
proj3/__init__.py:8-12 `resize_input` calls) against tests/golden/resize.npz, filled by
oracle/make_golden.py with cv2's own outputs; and against cv2 directly when it is importable."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
import resize as rs

G = np.load(os.path.join(GOLDEN, "resize.npz"))


def test_restatement_reproduces_cv2_goldens():
    for i, (h, w, d) in enumerate(G["cases"]):
        got = rs.resize_bilinear_u8(G["src_%d" % i], int(d), int(d))
        np.testing.assert_array_equal(got, G["dst_%d" % i], err_msg="case %d (%dx%d -> %d)" % (i, h, w, d))
    big = rs.resize_bilinear_u8(G["big_src"], 416, 416)
    assert int(big.astype(np.int64).sum()) == int(G["big_sum"])
    np.testing.assert_array_equal(big[::13, ::7], G["big_sample"])


def test_against_cv2_itself_when_present():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for h, w in [(480, 640), (375, 500), (208, 208), (833, 831), (2, 2), (720, 1280)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        np.testing.assert_array_equal(rs.resize_bilinear_u8(img, 416, 416), cv2.resize(img, (416, 416)))


def test_resize_input_is_the_reference_chain():
    frame = G["src_0"]
    x = rs.resize_input(frame, 96)
    assert x.dtype == np.float32 and x.shape == (96, 96, 3)
    want = np.asarray((G["dst_0"] / 255.)[:, :, ::-1], dtype=np.float32)
    np.testing.assert_array_equal(x, want)


def test_identity_size_is_a_copy():
    np.testing.assert_array_equal(rs.resize_bilinear_u8(G["src_2"], 96, 96), G["src_2"])
