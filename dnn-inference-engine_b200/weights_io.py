"""Weight containers for the YOLOv2-tiny path ("next" row f2).

The reference keeps its weights in ``y2t_weights.pickle``: a list of 9 dicts with keys ``kernel``
(HWIO), ``biases`` and, for layers 0-7, ``moving_mean``, ``moving_variance``, ``gamma``
(proj3/yolov2tiny.py:15-23,30-77; written by a Python-2 TensorFlow script, hence the latin1 decode).
The pickle itself is not in the reference repository, so this module also reads and writes the
format those numbers originally come from: darknet's ``yolov2-tiny-voc.weights``.
"""
import pickle
import struct

import numpy as np

KEYS_BN = ("kernel", "biases", "moving_mean", "moving_variance", "gamma")


def _validate(layers):
    if not isinstance(layers, (list, tuple)) or not layers:
        raise ValueError("expected a non-empty list of per-layer dicts")
    out = []
    for i, w in enumerate(layers):
        if "kernel" not in w or "biases" not in w:
            raise ValueError("layer %d lacks kernel/biases" % i)
        d = {k: np.ascontiguousarray(np.asarray(v), dtype=np.float32) for k, v in w.items() if k in KEYS_BN}
        if d["kernel"].ndim != 4 or d["biases"].shape != (d["kernel"].shape[3],):
            raise ValueError("layer %d: kernel must be HWIO and biases [O]" % i)
        bn = [k in d for k in ("moving_mean", "moving_variance", "gamma")]
        if any(bn) and not all(bn):
            raise ValueError("layer %d: batch norm needs moving_mean, moving_variance and gamma" % i)
        if i > 0 and d["kernel"].shape[2] != out[-1]["kernel"].shape[3]:
            raise ValueError("layer %d: %d input channels after %d output channels" %
                             (i, d["kernel"].shape[2], out[-1]["kernel"].shape[3]))
        out.append(d)
    return out


def load_pickle(path):
    """The reference's own container (proj3/yolov2tiny.py:15-23), any Python version; everything is
    returned as float32 (the reference passes the arrays to C as float* without casting)."""
    with open(path, "rb") as h:
        try:
            layers = pickle.load(h)
        except UnicodeDecodeError:
            h.seek(0)
            layers = pickle.load(h, encoding="latin1")
    return _validate(layers)


def save_pickle(path, layers, protocol=2):
    with open(path, "wb") as h:
        pickle.dump([dict(w) for w in _validate(layers)], h, protocol=protocol)


def load_darknet(path, channels, ksizes, batch_norm=None):
    """darknet .weights: header (major, minor, revision int32; `seen` int64 when major*10+minor >= 2, else
    int32), then per conv layer: biases[O], and with batch norm scales[O] (= gamma), rolling_mean[O],
    rolling_variance[O]; then the filters as [O][I][H][W] float32.  Returned in the reference's layout (HWIO)."""
    n = len(ksizes)
    batch_norm = [i < n - 1 for i in range(n)] if batch_norm is None else batch_norm
    with open(path, "rb") as h:
        major, minor, _rev = struct.unpack("<3i", h.read(12))
        h.read(8 if major * 10 + minor >= 2 else 4)
        data = np.frombuffer(h.read(), dtype="<f4")
    pos, layers = 0, []

    def take(count):
        nonlocal pos
        if pos + count > data.size:
            raise ValueError("darknet file too short for this architecture")
        out = data[pos:pos + count]
        pos += count
        return np.array(out, dtype=np.float32)

    for i in range(n):
        k, ic, oc = ksizes[i], channels[i], channels[i + 1]
        w = {"biases": take(oc)}
        if batch_norm[i]:
            w["gamma"], w["moving_mean"], w["moving_variance"] = take(oc), take(oc), take(oc)
        w["kernel"] = np.ascontiguousarray(take(oc * ic * k * k).reshape(oc, ic, k, k).transpose(2, 3, 1, 0))
        layers.append(w)
    if pos != data.size:
        raise ValueError("darknet file has %d floats left over for this architecture" % (data.size - pos))
    return _validate(layers)


def save_darknet(path, layers, seen=0):
    layers = _validate(layers)
    with open(path, "wb") as h:
        h.write(struct.pack("<3i", 0, 2, 0))
        h.write(struct.pack("<q", seen))
        for w in layers:
            h.write(w["biases"].astype("<f4").tobytes())
            if "gamma" in w:
                for key in ("gamma", "moving_mean", "moving_variance"):
                    h.write(w[key].astype("<f4").tobytes())
            h.write(np.ascontiguousarray(w["kernel"].transpose(3, 2, 0, 1)).astype("<f4").tobytes())
