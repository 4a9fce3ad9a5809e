"""GPU parity tests for the on-device detection decoding (csrc/postproc.cuh) against
oracle/postproc.py and the label boxes the reference's own `postprocessing` produced
(tests/golden/postproc.npz).

Tolerances: integer results (class, box edges, which boxes survive, their order) must be
IDENTICAL on every image whose decisions have a float64 margin (oracle `stable_images`); scores
within 2e-6 relative (numpy's float32 exp is 1-2 ulp from correctly rounded, the device rounds
correctly).  The NMS half alone, fed identical candidates, is exact with no exception."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
import _lib
import detect
import dnn_b200
import postproc as pp
from synth import build_yolov2tiny, make_images, make_weights, tame_head

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(GOLDEN, "postproc.npz"))
SCORE_RTOL = 2e-6


def golden_cases():
    seeded = pp.synth_predictions(int(G["seeded_batch"]), int(G["seeded_seed"]), float(G["seeded_spread"]))
    preds = np.concatenate([G["pred_stored"], seeded])
    ends = np.cumsum(G["det_count"])
    return preds, [G["det_rows"][e - c:e] for c, e in zip(G["det_count"], ends)]


def rows_of(dets, count):
    d = dets[:count]
    return np.stack([d["cls"], d["left"], d["top"], d["right"], d["bottom"]], axis=1).astype(np.int64).reshape(-1, 5)


def test_label_boxes_against_the_reference_function():
    preds, want = golden_cases()
    dets, counts = _lib.postprocess(preds, max_det=_lib.CANDIDATES)
    stable = pp.stable_images(preds)
    assert stable.mean() > 0.6
    exact = 0
    for i in range(len(preds)):
        got = rows_of(dets[i], counts[i])
        same = got.shape == want[i].shape and np.array_equal(got, want[i])
        exact += same
        if stable[i]:
            assert same, i
    assert exact >= len(preds) - 2
    # the drop-in function returns the reference's structure
    boxes = detect.postprocessing(preds[0])
    assert [b[0] for b in boxes] == [pp.CLASS_NAMES[c] for c in want[0][:, 0]]
    assert all(isinstance(b[1], tuple) and isinstance(b[2], tuple) and len(b[3]) == 3 for b in boxes)
    assert boxes == pp.postprocessing(preds[0]) or not stable[0]


def test_decoded_candidates_against_oracle():
    preds = pp.synth_predictions(12, seed=21, spread=1.2)
    cand = _lib.postprocess_candidates(preds)
    edges64, _, gap64 = pp.decode_f64(preds)
    n_edge_diff = n_cls_diff = 0
    for i, p in enumerate(preds):
        boxes, score, cls, _ = pp.decode_image(p)
        got = np.stack([cand[i]["left"], cand[i]["top"], cand[i]["right"], cand[i]["bottom"]], axis=1).astype(np.int64)
        np.testing.assert_allclose(cand[i]["score"], score, rtol=SCORE_RTOL, atol=1e-30)
        # an edge may differ only where the untruncated value sits within a few ulp of an integer
        diff = got != boxes
        near = np.abs(edges64[i] - np.round(edges64[i])) < 4e-6 * np.maximum(1.0, np.abs(edges64[i]))
        assert not np.any(diff & ~near)
        n_edge_diff += int(diff.sum())
        cdiff = cand[i]["cls"] != cls
        assert not np.any(cdiff & (gap64[i] > 1e-5))
        n_cls_diff += int(cdiff.sum())
    assert n_edge_diff <= 2 and n_cls_diff <= 1      # 12 * 845 * 4 edges


def test_nms_alone_is_exact_including_ties_and_disjoint_box_quirk():
    rng = np.random.default_rng(5)
    B = 24
    cand = np.zeros((B, _lib.CANDIDATES), dtype=_lib.DET_DTYPE)
    for i in range(B):
        n = int(rng.integers(0, 400)) if i else _lib.CANDIDATES          # image 0: every candidate passes
        cx, cy = rng.uniform(0, 416, _lib.CANDIDATES), rng.uniform(0, 416, _lib.CANDIDATES)
        w, h = rng.gamma(2.0, 40.0, _lib.CANDIDATES), rng.gamma(2.0, 40.0, _lib.CANDIDATES)
        cand[i]["left"], cand[i]["right"] = (cx - w / 2).astype(np.int32), (cx + w / 2).astype(np.int32)
        cand[i]["top"], cand[i]["bottom"] = (cy - h / 2).astype(np.int32), (cy + h / 2).astype(np.int32)
        cand[i]["cls"] = rng.integers(0, 20, _lib.CANDIDATES)
        s = np.round(rng.uniform(0.3, 1.0, _lib.CANDIDATES), 2 if i % 2 else 6).astype(np.float32)    # many exact score ties
        s[rng.permutation(_lib.CANDIDATES)[n:]] = 0.1
        cand[i]["score"] = s
    cand[1]["left"][:4], cand[1]["right"][:4] = [0, 0, 500, 0], [9, 9, 509, 9]         # duplicates + a far box
    cand[1]["top"][:4], cand[1]["bottom"][:4] = [0, 0, 500, 0], [9, 9, 509, 9]
    cand[1]["score"][:4] = [0.9, 0.9, 0.95, 0.5]
    # the pair on which the reference divides by zero (union == 0): the device suppresses the later box
    for j, (box, sc) in enumerate([([357, 31, 368, 75], 0.99), ([-15, 127, 48, 363], 0.98)]):
        cand[2]["left"][j], cand[2]["top"][j], cand[2]["right"][j], cand[2]["bottom"][j] = box
        cand[2]["score"][j] = sc
    dets, counts = _lib.nms(cand, max_det=_lib.CANDIDATES)
    thr = np.float32(0.3)
    for i in range(B):
        idx = np.flatnonzero(cand[i]["score"] > thr)
        boxes = np.stack([cand[i][k] for k in ("left", "top", "right", "bottom")], axis=1).astype(np.int64)
        kept = [idx[k] for k in pp.nms(boxes[idx], cand[i]["score"][idx], zero_union="drop")]
        assert counts[i] == len(kept)
        got = dets[i][:counts[i]]
        for name in ("cls", "left", "top", "right", "bottom", "score"):
            np.testing.assert_array_equal(got[name], cand[i][name][kept], err_msg="image %d %s" % (i, name))


def test_max_det_truncates_but_counts_everything():
    preds = pp.synth_predictions(4, seed=33)
    full, counts = _lib.postprocess(preds, max_det=_lib.CANDIDATES)
    cut, counts_cut = _lib.postprocess(preds, max_det=3)
    np.testing.assert_array_equal(counts, counts_cut)
    assert counts.max() > 3
    for i in range(4):
        np.testing.assert_array_equal(cut[i], full[i][:3])


def test_degenerate_inputs():
    # nothing passes; everything identical and confident; NaN / inf head values must not hang or crash
    quiet = np.full((1, 13, 13, 125), -30, np.float32)
    dets, counts = _lib.postprocess(quiet)
    assert counts[0] == 0
    same = np.zeros((1, 13, 13, 5, 25), np.float32)
    same[..., 4] = 8
    same[..., 5] = 12
    dets, counts = _lib.postprocess(same.reshape(1, 13, 13, 125), max_det=_lib.CANDIDATES)
    want = pp.detect_image(same[0])
    assert counts[0] == len(want)
    assert [tuple(int(dets[0][k][j]) for k in ("cls", "left", "top", "right", "bottom")) for j in range(counts[0])] == [w[:5] for w in want]
    wild = pp.synth_predictions(2, seed=3).copy()
    wild[0, 0, 0, :5] = [np.nan, np.inf, 200.0, -np.inf, 50.0]
    wild[1, 3, 3, 2:5] = [95.0, 95.0, 30.0]          # exp overflows float32: edges saturate at +-2^29
    wild[1, 3, 3, 5] = 40.0
    dets, counts = _lib.postprocess(wild, max_det=_lib.CANDIDATES)
    assert counts.min() >= 0 and counts.max() <= _lib.CANDIDATES
    assert np.abs(dets["left"]).max() <= 2 ** 29 and np.abs(dets["right"]).max() <= 2 ** 29
    with pytest.raises(RuntimeError):
        _lib.postprocess(quiet, max_det=0)


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_network_detect_equals_forward_then_oracle(precision):
    """Net.detect (forward + decoding on the device, boxes only over PCIe) against the oracle's
    decoding of the SAME head tensor, float32 and uint8-frame inputs, both halves of the pipeline."""
    B = 8
    weights = tame_head(make_weights(0))
    g = dnn_b200.DnnGraphBuilder()
    build_yolov2tiny(g, [B, 416, 416, 3], weights)
    eng = dnn_b200.DnnInferenceEngine(g, False, precision=precision)
    x = make_images(B, seed=300)
    head = eng.run(x).copy()
    net = eng._net
    net.set_detect(max_det=64)
    stable = pp.stable_images(head)
    assert stable.sum() >= 2
    n_boxes = 0
    for rep in range(2):
        dets, counts = net.detect(x)
        for i in range(B):
            want = pp.detect_image(head[i], zero_union="drop")
            n_boxes += len(want)
            if stable[i]:
                assert counts[i] == len(want)
                got = rows_of(dets[i], min(counts[i], 64))
                np.testing.assert_array_equal(got, np.array([w[:5] for w in want[:64]], dtype=np.int64).reshape(-1, 5))
                np.testing.assert_allclose(dets[i]["score"][:len(want[:64])], [w[5] for w in want[:64]], rtol=SCORE_RTOL)
    assert n_boxes > 0
    frames = np.random.default_rng(9).integers(0, 256, size=(B, 416, 416, 3), dtype=np.uint8)
    head_u8 = net.forward_u8(frames, bgr=True).copy()
    dets, counts = net.detect(frames, bgr=True)
    for i in np.flatnonzero(pp.stable_images(head_u8)):
        want = pp.detect_image(head_u8[i], zero_union="drop")
        assert counts[i] == len(want)
        np.testing.assert_array_equal(rows_of(dets[i], min(counts[i], 64)),
                                      np.array([w[:5] for w in want[:64]], dtype=np.int64).reshape(-1, 5))
    eng.close()


def test_detect_refuses_networks_without_a_yolo_head():
    g = dnn_b200.DnnGraphBuilder()
    i = g.create_input([1, 16, 16, 3])
    c = g.create_conv2d(i, np.zeros((3, 3, 3, 8), np.float32), [1, 1, 1, 1], 'SAME')
    g.set_out_node(g.create_bias_add(c, np.zeros(8, np.float32)))
    eng = dnn_b200.DnnInferenceEngine(g, False)
    eng.run(np.zeros((1, 16, 16, 3), np.float32))
    with pytest.raises(RuntimeError):
        eng._net.detect(np.zeros((1, 16, 16, 3), np.float32))
    eng.close()
