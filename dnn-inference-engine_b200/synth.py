"""Synthetic YOLOv2-tiny weights and inputs (there is no weights pickle in the reference
repo and no network here).  Recipe: SURVEY.md section 8(d).

The weight container is the one the reference consumes: a list of 9 dicts with keys
``kernel`` (HWIO), ``biases`` and, for layers 0-7, ``moving_mean``, ``moving_variance``,
``gamma`` (proj3/yolov2tiny.py:30-77).  Everything is float32: the reference passes the
arrays to C as ``float*`` without casting (proj3/dnn_openblas.py:166-167,197,251-253).
"""
import numpy as np

CHANNELS = [3, 16, 32, 64, 128, 256, 512, 1024, 1024, 125]
KSIZES = [3, 3, 3, 3, 3, 3, 3, 3, 1]
# pool after conv block i: (ksize, stride) or None      proj3/yolov2tiny.py:34,40,46,52,58,64
POOLS = [(2, 2), (2, 2), (2, 2), (2, 2), (2, 2), (2, 1), None, None, None]
EPSILON = 1e-5
IN_HW = 416


def make_weights(seed=0, channels=None, ksizes=None, negative_gamma=False):
    """Random-init weights, He-scaled so activations stay O(1) through the 9 layers."""
    channels = CHANNELS if channels is None else channels
    ksizes = KSIZES if ksizes is None else ksizes
    rng = np.random.default_rng(seed)
    layers = []
    n = len(ksizes)
    for i in range(n):
        k, ic, oc = ksizes[i], channels[i], channels[i + 1]
        w = {}
        std = np.sqrt(2.0 / (k * k * ic))
        w["kernel"] = (rng.standard_normal((k, k, ic, oc)) * std).astype(np.float32)
        w["biases"] = (0.1 * rng.standard_normal(oc)).astype(np.float32)
        if i < n - 1:
            w["moving_mean"] = (0.1 * rng.standard_normal(oc)).astype(np.float32)
            w["moving_variance"] = rng.uniform(0.5, 1.5, oc).astype(np.float32)
            g = rng.uniform(0.5, 1.5, oc).astype(np.float32)
            if negative_gamma:
                g[::3] *= -1.0
            w["gamma"] = g
        layers.append(w)
    return layers


def make_images(batch, seed=100, hw=IN_HW, start=0):
    """Images start..start+batch-1 of the synthetic stream: uniform [0,1) float32 NHWC, the
    range proj3/__init__.py:8-12 (resize_input) produces.  Image i depends only on seed+i,
    so any shard of a big batch can be regenerated for a parity check."""
    out = np.empty((batch, hw, hw, 3), dtype=np.float32)
    for i in range(batch):
        out[i] = np.random.default_rng(seed + start + i).random((hw, hw, 3), dtype=np.float32)
    return out


def build_yolov2tiny(builder, in_shape, weights, copy_variance=True):
    """Wire the proj3 YOLOv2-tiny layer stack (proj3/yolov2tiny.py:28-79) on any builder that
    has the reference's DnnGraphBuilder interface.  Returns the list of created nodes in
    execution order (input first).  ``copy_variance``: the reference's batch_norm overwrites
    the array it is handed (proj3/dnn_openblas.c:48-50), so give it a private copy."""
    g = builder
    nodes = [g.create_input(in_shape)]
    cur = nodes[0]
    n = len(weights)
    for i, w in enumerate(weights):
        cur = g.create_conv2d(cur, w["kernel"], strides=[1, 1, 1, 1], padding='SAME'); nodes.append(cur)
        cur = g.create_bias_add(cur, w["biases"]); nodes.append(cur)
        if i < n - 1:
            var = w["moving_variance"].copy() if copy_variance else w["moving_variance"]
            cur = g.create_batch_norm(cur, w["moving_mean"], var, w["gamma"], EPSILON); nodes.append(cur)
            cur = g.create_leaky_relu(cur); nodes.append(cur)
            pool = POOLS[i] if n == len(POOLS) else None
            if pool is not None:
                k, s = pool
                cur = g.create_max_pool2d(cur, ksize=[1, k, k, 1], strides=[1, s, s, 1], padding='SAME')
                nodes.append(cur)
    g.set_out_node(cur)
    return nodes


def tame_head(weights, objectness_shift=-2.5):
    """Random-init weights give head outputs of std ~4.5, i.e. boxes e^20 pixels wide on which
    the reference's own `postprocessing` divides by zero.  For detection tests rescale the last
    layer per field so that the head looks like a detector's: box offsets and sizes of order one,
    objectness mostly low, class logits peaked -- a handful of confident boxes per image."""
    field_scale = np.array([0.3, 0.3, 0.1, 0.1, 0.4] + [0.45] * 20, dtype=np.float32)
    col_scale = np.tile(field_scale, 5)
    out = [dict(w) for w in weights]
    last = dict(out[-1])
    last["kernel"] = (last["kernel"] * col_scale).astype(np.float32)
    b = (last["biases"] * col_scale).astype(np.float32)
    b[4::25] += objectness_shift
    last["biases"] = b
    out[-1] = last
    return out
