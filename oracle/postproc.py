"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's detection head decoding
(`postprocessing`, `iou`, `non_maximal_suppression`, `sigmoid`, `softmax`:
cs492-projects/proj3/yolov2tiny.py:94-235).  Only tests/, __graft_entry__.smoke() and bench.py's
baseline legs may import it; the product path (csrc/postproc.cuh) never does.

Pinned: oracle/make_golden.py runs the reference's own `postprocessing` on seeded prediction
tensors and stores its label boxes in tests/golden/postproc.npz; tests/test_oracle.py checks this
restatement against them.

What the reference computes, per image, on the [13,13,125] head output viewed as
[row, col, anchor box b, (tx, ty, tw, th, tc, 20 class logits)]  (yolov2tiny.py:122):
  sigmoid(x)  = 1 / (1 + np.e ** -x)                                   (:228-229)
  center_x    = (col + sigmoid(tx)) * 32,   center_y = (row + sigmoid(ty)) * 32     (:134-135)
  roi_w       = exp(tw) * anchor_w[b] * 32, roi_h = exp(th) * anchor_h[b] * 32      (:137-138)
  score       = sigmoid(tc) * max(softmax(class logits)); class = FIRST arg-max     (:140-148,156)
  box         = int(center -/+ roi/2), truncation toward zero                       (:151-154)
  keep candidates with score > 0.3 in (row, col, b) raster order                    (:156-157)
  stable sort by score, descending                                                  (:160)
  greedy class-agnostic NMS: a candidate is dropped when its IoU with ANY already kept
  box exceeds 0.3                                                                   (:204-226)
  IoU uses the +1 pixel convention and does NOT clamp an empty intersection to zero: two boxes
  disjoint on both axes get a positive "intersection" (product of two negative extents) (:182-198)

Arithmetic type: the inputs are numpy float32 scalars; with NumPy >= 2 (this image: 2.3) Python
floats are weak, so every step above stays float32 (np.e is rounded to float32 before the power).
The restatement performs the same scalar numpy operations in the same order, so on one machine
it reproduces the reference bit for bit.  Across machines the last ulp of numpy's float32
exp/pow may differ (SIMD dispatch), which is why comparisons go through `stable_images`.
"""
import numpy as np

GRID = 13
BOXES = 5
CLASSES = 20
CAND = GRID * GRID * BOXES
ANCHORS = (1.08, 1.19, 3.42, 4.41, 6.63, 11.38, 9.42, 5.11, 16.62, 10.52)      # yolov2tiny.py:117
CLASS_NAMES = ("aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow",
               "diningtable", "dog", "horse", "motorbike", "person", "pottedplant", "sheep", "sofa",
               "train", "tvmonitor")                                           # Pascal VOC order, :103
SCORE_THRESHOLD = 0.3   # :156
IOU_THRESHOLD = 0.3     # :164


def class_color(i):
    """The reference's colour table (:104-114) is an arithmetic pattern over the class index
    (checked value for value against the table in tests/golden/postproc.npz)."""
    return (254 * (1 - i / 18), 254 * (1 - (i % 9) / 6), (254, 127, 0)[i % 3])


def _sigmoid(x):
    return 1 / (1 + np.e ** -x)


def decode_image(pred, score_threshold=SCORE_THRESHOLD):
    """All 845 candidates of one image, in raster order.
    Returns (boxes int64 [845,4] = left, top, right, bottom; score float32 [845]; cls int32 [845];
    passed bool [845])."""
    p = np.reshape(np.asarray(pred, dtype=np.float32), (GRID, GRID, BOXES, 5 + CLASSES))
    boxes = np.zeros((CAND, 4), dtype=np.int64)
    score = np.zeros(CAND, dtype=np.float32)
    cls = np.zeros(CAND, dtype=np.int32)
    passed = np.zeros(CAND, dtype=bool)
    i = 0
    for row in range(GRID):
        for col in range(GRID):
            for b in range(BOXES):
                tx, ty, tw, th, tc = p[row, col, b, :5]
                cx = (float(col) + _sigmoid(tx)) * 32.0
                cy = (float(row) + _sigmoid(ty)) * 32.0
                w = np.exp(tw) * ANCHORS[2 * b] * 32.0
                h = np.exp(th) * ANCHORS[2 * b + 1] * 32.0
                conf = _sigmoid(tc)
                logits = p[row, col, b, 5:]
                e = np.exp(logits - np.max(logits))
                prob = tuple(e / e.sum(axis=0))
                best = prob.index(max(prob))
                s = conf * prob[best]
                boxes[i] = (int(cx - (w / 2.)), int(cy - (h / 2.)), int(cx + (w / 2.)), int(cy + (h / 2.)))
                score[i] = s
                cls[i] = best
                passed[i] = s > score_threshold
                i += 1
    return boxes, score, cls, passed


def iou(a, b):
    """+1 pixel convention, unclamped intersection (:182-198); Python ints, true division."""
    xa, ya = max(a[0], b[0]), max(a[1], b[1])
    xb, yb = min(a[2], b[2]), min(a[3], b[3])
    inter = (xb - xa + 1) * (yb - ya + 1)
    area_a = (a[2] - a[0] + 1) * (a[3] - a[1] + 1)
    area_b = (b[2] - b[0] + 1) * (b[3] - b[1] + 1)
    return inter / float(area_a + area_b - inter)


def nms(boxes, score, iou_threshold=IOU_THRESHOLD, zero_union="raise"):
    """boxes [n,4] ints and score [n] of the thresholded candidates in raster order.
    Returns the indices kept, in the order the reference emits them (score descending, stable).
    A pair whose union is zero (two boxes disjoint on both axes whose unclamped "intersection"
    equals the sum of their areas, e.g. [357,31,368,75] and [-15,127,48,363]) makes the reference
    raise ZeroDivisionError (:196); zero_union="drop" is the device's documented behaviour
    instead: the IEEE quotient is +inf, the candidate is suppressed."""
    order = sorted(range(len(score)), key=lambda i: score[i], reverse=True)
    kept = []
    for i in order:
        bi = [int(v) for v in boxes[i]]
        drop = False
        for j in kept:
            try:
                over = iou(bi, [int(v) for v in boxes[j]]) > iou_threshold
            except ZeroDivisionError:
                if zero_union != "drop":
                    raise
                over = True
            if over:
                drop = True
        if not drop:
            kept.append(i)
    return kept


def detect_image(pred, score_threshold=SCORE_THRESHOLD, iou_threshold=IOU_THRESHOLD, zero_union="raise"):
    """[(cls, left, top, right, bottom, score)] of one image after threshold + NMS."""
    boxes, score, cls, passed = decode_image(pred, score_threshold)
    idx = np.flatnonzero(passed)
    kept = nms(boxes[idx], score[idx], iou_threshold, zero_union)
    return [(int(cls[idx[k]]),) + tuple(int(v) for v in boxes[idx[k]]) + (float(score[idx[k]]),) for k in kept]


def postprocessing(pred):
    """Same return value as the reference function: [(class name, (left, top), (right, bottom), colour)]."""
    return [(CLASS_NAMES[c], (l, t), (r, b), class_color(c)) for c, l, t, r, b, _ in detect_image(pred)]


# ------------------------------------------------------------------------------------------------
# float64 margins: how far an image is from every discrete decision the pipeline takes.  Two
# float32 implementations whose transcendental functions differ in the last ulp must agree exactly
# on images whose margins are all comfortably larger than that.
def decode_f64(pred):
    p = np.asarray(pred, dtype=np.float64).reshape(-1, GRID, GRID, BOXES, 5 + CLASSES)
    col = np.arange(GRID, dtype=np.float64)[None, None, :, None]
    row = np.arange(GRID, dtype=np.float64)[None, :, None, None]
    aw = np.array(ANCHORS[0::2])[None, None, None, :]
    ah = np.array(ANCHORS[1::2])[None, None, None, :]
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    with np.errstate(over="ignore"):
        cx, cy = (col + sig(p[..., 0])) * 32.0, (row + sig(p[..., 1])) * 32.0
        w, h = np.exp(p[..., 2]) * aw * 32.0, np.exp(p[..., 3]) * ah * 32.0
        logits = p[..., 5:]
        e = np.exp(logits - logits.max(-1, keepdims=True))
    prob = e / e.sum(-1, keepdims=True)
    srt = np.sort(prob, axis=-1)
    score = sig(p[..., 4]) * srt[..., -1]
    edges = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=-1)
    n = p.shape[0]
    return edges.reshape(n, CAND, 4), score.reshape(n, CAND), (srt[..., -1] - srt[..., -2]).reshape(n, CAND)


def stable_images(pred, score_threshold=SCORE_THRESHOLD, iou_threshold=IOU_THRESHOLD, rel=2e-6, max_coord=2.0 ** 29):
    """bool [B]: images whose decisions all have a margin above `rel` (relative) in float64:
    the candidates' distance from the score threshold, every near-threshold candidate's box edges
    from the next integer, its top-2 class gap, gaps between sorted scores, every pair's IoU from
    the IoU threshold; and all coordinates within the device's exact range."""
    edges, score, gap = decode_f64(pred)
    ok = np.ones(edges.shape[0], dtype=bool)
    for n in range(edges.shape[0]):
        s = score[n]
        if np.any(np.abs(s - score_threshold) < rel * 4):
            ok[n] = False
            continue
        idx = np.flatnonzero(s > score_threshold)
        if idx.size == 0:
            continue
        e = edges[n, idx]
        if not np.all(np.isfinite(e)) or np.abs(e).max() >= max_coord:
            ok[n] = False
            continue
        frac = np.abs(e - np.round(e))
        if np.any(frac < 0.5 * rel * np.maximum(1.0, np.abs(e))) or np.any(gap[n, idx] < rel * 8):
            ok[n] = False
            continue
        ss = np.sort(s[idx])
        if ss.size > 1 and np.any(np.diff(ss) < rel * 4 * ss[1:]):
            ok[n] = False
            continue
        b = np.trunc(e).astype(np.int64)
        xa = np.maximum(b[:, None, 0], b[None, :, 0]); ya = np.maximum(b[:, None, 1], b[None, :, 1])
        xb = np.minimum(b[:, None, 2], b[None, :, 2]); yb = np.minimum(b[:, None, 3], b[None, :, 3])
        inter = (xb - xa + 1).astype(np.float64) * (yb - ya + 1)
        area = (b[:, 2] - b[:, 0] + 1).astype(np.float64) * (b[:, 3] - b[:, 1] + 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            io = inter / (area[:, None] + area[None, :] - inter)
        if np.any(~np.isfinite(io)) or np.any(np.abs(io - iou_threshold) < 1e-9):
            ok[n] = False
    return ok


def synth_predictions(batch, seed, spread=1.0):
    """Head outputs with the statistics of a trained detector rather than of random weights:
    mostly background (objectness logit well below zero) with a few confident, peaked cells."""
    rng = np.random.default_rng(seed)
    p = np.empty((batch, GRID, GRID, BOXES, 5 + CLASSES), dtype=np.float32)
    p[..., 0:2] = rng.normal(0.0, 1.5, p[..., 0:2].shape)
    p[..., 2:4] = rng.normal(-0.3, 0.7 * spread, p[..., 2:4].shape)
    p[..., 4] = rng.normal(-4.0, 2.5, p[..., 4].shape)
    p[..., 5:] = rng.normal(0.0, 1.0, p[..., 5:].shape)
    hot = rng.random(p[..., 4].shape) < 0.04            # cells that "see" an object
    p[..., 4] = np.where(hot, rng.normal(2.0, 1.5, hot.shape), p[..., 4])
    peak = rng.integers(0, CLASSES, hot.shape)
    boost = np.where(hot, rng.normal(6.0, 2.0, hot.shape), 0.0).astype(np.float32)
    np.put_along_axis(p[..., 5:], peak[..., None], np.take_along_axis(p[..., 5:], peak[..., None], -1) + boost[..., None], -1)
    return p.reshape(batch, GRID, GRID, BOXES * (5 + CLASSES))
