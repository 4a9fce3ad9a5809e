"""detect -- the step after the layer stack, on the GPU: drop-in for `postprocessing` of
proj3/yolov2tiny.py:94-180 (sigmoid / exp / softmax decoding of the [13,13,125] head tensor,
score threshold 0.3, stable sort, greedy class-agnostic NMS at IoU 0.3).

    from detect import postprocessing            # same argument, same return value
    boxes = postprocessing(np.squeeze(tout))     # [(class name, (left, top), (right, bottom), colour)]

`postprocessing_batch` does a whole batch in one kernel launch, and `Net.detect` (_lib.py) chains it
behind the network so that only the boxes cross PCIe.  The arithmetic runs in
csrc/postproc.cuh; this module only maps class indices to the reference's names and colours
(proj3/yolov2tiny.py:103-114).  No CPU path.
"""
import numpy as np

import _lib

CLASS_NAMES = ("aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow",
               "diningtable", "dog", "horse", "motorbike", "person", "pottedplant", "sheep", "sofa",
               "train", "tvmonitor")


def class_color(i):
    """Colour of class i: the reference's 20-entry table follows this pattern exactly."""
    return (254 * (1 - i / 18), 254 * (1 - (i % 9) / 6), (254, 127, 0)[i % 3])


def label_boxes(dets, count):
    """One image's rows of _lib.DET_DTYPE -> the reference's list of label boxes."""
    out = []
    for d in dets[:min(int(count), len(dets))]:
        c = int(d["cls"])
        out.append((CLASS_NAMES[c], (int(d["left"]), int(d["top"])), (int(d["right"]), int(d["bottom"])), class_color(c)))
    return out


def postprocessing(predictions, device=0):
    """predictions: the head output of ONE image, any shape with 13*13*125 elements."""
    pred = np.reshape(np.asarray(predictions, dtype=np.float32), (1, 13, 13, 125))
    dets, counts = _lib.postprocess(pred, max_det=_lib.CANDIDATES, device=device)
    return label_boxes(dets[0], counts[0])


def postprocessing_batch(predictions, max_det=64, device=0):
    """predictions [B,13,13,125] -> list of B label-box lists (at most max_det boxes each)."""
    dets, counts = _lib.postprocess(predictions, max_det=max_det, device=device)
    return [label_boxes(d, c) for d, c in zip(dets, counts)]
