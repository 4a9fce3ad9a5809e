"""CPU tests: oracle/postproc.py (restatement of the reference's `postprocessing`,
proj3/yolov2tiny.py:94-235) against tests/golden/postproc.npz, which oracle/make_golden.py
filled by calling the reference function itself."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
import postproc as pp

G = np.load(os.path.join(GOLDEN, "postproc.npz"))


def golden_cases():
    """[(prediction [13,13,125], reference rows [n,5] = class, left, top, right, bottom)]"""
    seeded = pp.synth_predictions(int(G["seeded_batch"]), int(G["seeded_seed"]), float(G["seeded_spread"]))
    assert abs(seeded.astype(np.float64).sum() - float(G["seeded_checksum"])) < 1e-6, "seeded predictions drifted"
    preds = list(G["pred_stored"]) + list(seeded)
    ends = np.cumsum(G["det_count"])
    return [(p, G["det_rows"][e - c:e]) for p, c, e in zip(preds, G["det_count"], ends)]


def test_restatement_reproduces_the_reference_function():
    cases = golden_cases()
    stable = pp.stable_images(np.stack([p for p, _ in cases]))
    assert stable.mean() > 0.6
    exact = 0
    for (p, rows), ok in zip(cases, stable):
        got = np.array([d[:5] for d in pp.detect_image(p)], dtype=np.int64).reshape(-1, 5)
        same = got.shape == rows.shape and np.array_equal(got, rows)
        exact += same
        if ok:      # every decision has a margin far above an ulp: must be identical on any machine
            assert same
    # the same numpy scalar operations in the same order: in practice all of them agree
    assert exact >= len(cases) - 2


def test_return_value_has_the_reference_shape():
    p = np.full((13, 13, 5, 25), -20, np.float32)
    p[6, 6, 0, :5] = [0, 0, 0, 0, 10]
    p[6, 6, 0, 5] = 20
    (name, lt, rb, color), = pp.postprocessing(p.reshape(13, 13, 125))
    assert name == "aeroplane" and isinstance(lt, tuple) and isinstance(rb, tuple)
    assert [lt[0], lt[1], rb[0], rb[1]] == list(G["single_box"])
    for c in range(pp.CLASSES):
        assert tuple(G["colors"][c]) == tuple(float(v) for v in pp.class_color(c))


def test_empty_image_yields_no_boxes():
    assert pp.postprocessing(np.full((13, 13, 125), -30, np.float32)) == []


def test_iou_does_not_clamp_disjoint_boxes():
    # disjoint on both axes: (negative) x (negative) = positive "intersection", as in the reference
    a, b = [0, 0, 9, 9], [20, 20, 29, 29]
    assert pp.iou(a, b) == 100 / float(100 + 100 - 100)
    # disjoint on one axis only: negative
    assert pp.iou([0, 0, 9, 9], [20, 0, 29, 9]) < 0


def test_nms_is_class_agnostic_and_order_stable():
    boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [100, 0, 110, 10], [1, 1, 11, 11]])
    score = np.array([0.5, 0.5, 0.9, 0.8], dtype=np.float32)
    # order: 2 (0.9), 3 (0.8), 0 (0.5, first of the tie), 1; 0 and 1 overlap 3
    assert pp.nms(boxes, score) == [2, 3]
    assert pp.nms(boxes[:2], score[:2]) == [0]


def test_zero_union_pair_raises_like_the_reference_unless_told_to_drop():
    boxes = np.array([[357, 31, 368, 75], [-15, 127, 48, 363]])
    score = np.array([0.9, 0.8], dtype=np.float32)
    with pytest.raises(ZeroDivisionError):
        pp.nms(boxes, score)
    assert pp.nms(boxes, score, zero_union="drop") == [0]
