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


DARKNET_EPS = 1e-6       # darknet normalize_cpu: (x - mean) / (sqrt(variance) + .000001f)
REFERENCE_EPS = 1e-5     # proj3/yolov2tiny.py:32 passes 1e-5 to create_batch_norm
_ZERO_SCALE_GAMMA = 1e-20


def darknet_to_reference(biases, scales, mean, variance, eps=REFERENCE_EPS):
    """One batch-normalised darknet layer -> the reference's (biases, gamma, moving_mean, moving_variance).

    The two formats apply the same four vectors differently:
      darknet   : y = scales * (conv - mean) / (sqrt(variance) + 1e-6) + biases          (bias AFTER the scale)
      reference : y = gamma * (conv + biases - mean) / sqrt(moving_variance + eps)        (proj3/yolov2tiny.py:30-33,
                                                                                          dnn_openblas.c:9-65)
    so copying them across gives a different network.  This keeps y identical: gamma = scales, mean unchanged,
    moving_variance = (sqrt(variance) + 1e-6)^2 - eps, biases_ref = biases * (sqrt(variance) + 1e-6) / scales.
    A channel with scales == 0 is a constant (y = biases) the reference form cannot express with gamma = 0; it gets
    gamma = 1e-20, whose conv term vanishes in float32 next to the (finite) bias term that carries the constant."""
    b, s, m, v = (np.asarray(a, dtype=np.float64) for a in (biases, scales, mean, variance))
    sd = np.sqrt(v) + DARKNET_EPS
    gamma = np.where(s == 0.0, _ZERO_SCALE_GAMMA, s)
    bias_ref = b * sd / gamma
    if not np.all(np.abs(bias_ref) < 3e38):
        raise ValueError("darknet layer cannot be expressed in the reference's bias-before-batch-norm form")
    return (bias_ref.astype(np.float32), gamma.astype(np.float32), m.astype(np.float32),
            (sd * sd - eps).astype(np.float32))


def reference_to_darknet(biases, gamma, mean, variance, eps=REFERENCE_EPS):
    """Inverse of darknet_to_reference: -> (biases, scales, rolling_mean, rolling_variance)."""
    b, g, m, v = (np.asarray(a, dtype=np.float64) for a in (biases, gamma, mean, variance))
    sd = np.sqrt(v + eps)
    if np.any(sd <= DARKNET_EPS):
        raise ValueError("moving_variance + eps too small for darknet's sqrt(variance) + 1e-6 form")
    return ((b * g / sd).astype(np.float32), g.astype(np.float32), m.astype(np.float32),
            ((sd - DARKNET_EPS) ** 2).astype(np.float32))


def load_darknet(path, channels, ksizes, batch_norm=None, convert=True, eps=REFERENCE_EPS):
    """darknet .weights: header (major, minor, revision int32; `seen` int64 when major*10+minor >= 2, else
    int32), then per conv layer: biases[O], and with batch norm scales[O], rolling_mean[O],
    rolling_variance[O]; then the filters as [O][I][H][W] float32.  Returned in the reference's layout (HWIO)
    and, with convert=True (default), in the reference's batch-norm convention (darknet_to_reference), so that
    the graph of proj3/yolov2tiny.py:25-79 computes what darknet computes.  convert=False returns darknet's
    vectors under the reference's key names unchanged (a different network: only for inspecting files)."""
    n = len(ksizes)
    batch_norm = [i < n - 1 for i in range(n)] if batch_norm is None else batch_norm
    with open(path, "rb") as h:
        head = h.read(12)
        if len(head) < 12:
            raise ValueError("darknet file too short for a header")
        major, minor, _rev = struct.unpack("<3i", head)
        seen_bytes = 8 if major * 10 + minor >= 2 else 4
        if len(h.read(seen_bytes)) < seen_bytes:
            raise ValueError("darknet file too short for a header")
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
            scales, mean, var = take(oc), take(oc), take(oc)
            if convert:
                w["biases"], w["gamma"], w["moving_mean"], w["moving_variance"] = darknet_to_reference(
                    w["biases"], scales, mean, var, eps)
            else:
                w["gamma"], w["moving_mean"], w["moving_variance"] = scales, mean, var
        w["kernel"] = np.ascontiguousarray(take(oc * ic * k * k).reshape(oc, ic, k, k).transpose(2, 3, 1, 0))
        layers.append(w)
    if pos != data.size:
        raise ValueError("darknet file has %d floats left over for this architecture" % (data.size - pos))
    return _validate(layers)


def save_darknet(path, layers, seen=0, version=(0, 2, 0), convert=True, eps=REFERENCE_EPS):
    """Writes reference-convention weights as a darknet file (inverse of load_darknet).  version (0, 1, x) writes
    the old 4-byte `seen` header."""
    layers = _validate(layers)
    major, minor, rev = version
    with open(path, "wb") as h:
        h.write(struct.pack("<3i", major, minor, rev))
        h.write(struct.pack("<q" if major * 10 + minor >= 2 else "<i", seen))
        for w in layers:
            if "gamma" in w:
                vecs = (reference_to_darknet(w["biases"], w["gamma"], w["moving_mean"], w["moving_variance"], eps)
                        if convert else (w["biases"], w["gamma"], w["moving_mean"], w["moving_variance"]))
                for v in vecs:
                    h.write(np.asarray(v).astype("<f4").tobytes())
            else:
                h.write(w["biases"].astype("<f4").tobytes())
            h.write(np.ascontiguousarray(w["kernel"].transpose(3, 2, 0, 1)).astype("<f4").tobytes())
