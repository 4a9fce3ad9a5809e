"""TEST INFRASTRUCTURE ONLY -- host glue that drives a CPU checker library through the
YOLOv2-tiny op chain the way the reference's host layer does (np.pad, then the C op).

Three backends share one builder interface (the reference's DnnGraphBuilder.create_* names,
so dnn-inference-engine_b200/synth.build_yolov2tiny can wire any of them):

  "oracle"        oracle/liboracle.so, our C restatement (any batch size)
  "ref_openblas"  oracle/_ref/libdnn_openblas.so = proj3/dnn_openblas.c compiled unmodified,
                  called with the argument lists of proj3/dnn_openblas.py:180-188,203-209,
                  236-242,262-270,281-284.  Batch 1 only, variance is copied
                  per call because the C code overwrites it (App. B1).
  "ref_avx"       oracle/_ref/libdnn_avx.so = proj3/dnn_avx.c compiled unmodified, called with
                  the argument lists of proj3/dnn_avx.py:202-207,234-238,286-289,316-321,330-333.
                  Batch 1 only; result buffers are zeroed per call (App. B4).

/root/reference is not needed at run time: only the prebuilt .so files are.
"""
import ctypes
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FP = ctypes.POINTER(ctypes.c_float)
IP = ctypes.POINTER(ctypes.c_int)
F32_MIN = np.finfo(np.float32).min

_libs = {}


def load(backend):
    if backend not in _libs:
        path = {"oracle": os.path.join(HERE, "liboracle.so"),
                "ref_openblas": os.path.join(HERE, "_ref", "libdnn_openblas.so"),
                "ref_avx": os.path.join(HERE, "_ref", "libdnn_avx.so")}[backend]
        if not os.path.exists(path):
            raise FileNotFoundError(path + " missing: run python oracle/build.py")
        _libs[backend] = ctypes.CDLL(path)
    return _libs[backend]


def available(backend):
    try:
        load(backend)
        return True
    except (FileNotFoundError, OSError):
        return False


def fp(a):
    return a.ctypes.data_as(FP)


def out_pads(in_size, k, s, padding):
    """proj3/dnn_openblas.py:127-142."""
    if padding == 'SAME':
        o = math.ceil(in_size / s)
        pad = max((o - 1) * s + k - in_size, 0)
        return o, pad // 2, pad - pad // 2
    if padding == 'VALID':
        return math.ceil((in_size - k + 1) / s), 0, 0
    raise ValueError(padding)


class _Node:
    def __init__(self, kind, in_node):
        self.kind, self.in_node, self.result = kind, in_node, None


class Chain:
    """Sequential op chain with the reference builder's method names."""

    def __init__(self, backend="oracle"):
        self.backend = backend
        self.lib = load(backend)
        self.nodes = []
        self.out_node = None
        self.in_shape = None

    # -- builder interface ------------------------------------------------------------
    def create_input(self, in_shape):
        self.in_shape = tuple(in_shape)
        n = _Node("input", None); n.shape = self.in_shape
        self.nodes.append(n)
        return n

    def _add(self, n, shape):
        n.shape = tuple(shape)
        self.nodes.append(n)
        return n

    def create_conv2d(self, in_node, kernel, strides, padding):
        b, h, w, c = in_node.shape
        kh, kw, kc, od = kernel.shape
        if kc != c:
            raise ValueError
        n = _Node("conv2d", in_node)
        n.kernel = np.ascontiguousarray(kernel, dtype=np.float32)
        n.strides = strides
        n.oh, n.pt, n.pb = out_pads(h, kh, strides[1], padding)
        n.ow, n.pl, n.pr = out_pads(w, kw, strides[2], padding)
        n.kernel_r = np.ascontiguousarray(n.kernel.transpose(2, 0, 1, 3).reshape(-1, od))
        return self._add(n, (b, n.oh, n.ow, od))

    def create_bias_add(self, in_node, biases):
        if not (biases.ndim == 1 and in_node.shape[-1] == biases.shape[0]):
            raise ValueError
        n = _Node("bias_add", in_node)
        n.biases = np.ascontiguousarray(biases, dtype=np.float32)
        return self._add(n, in_node.shape)

    def create_batch_norm(self, in_node, mean, variance, gamma, epsilon):
        n = _Node("batch_norm", in_node)
        n.mean = np.ascontiguousarray(mean, dtype=np.float32)
        n.variance = np.array(variance, dtype=np.float32, copy=True)
        n.gamma = np.ascontiguousarray(gamma, dtype=np.float32)
        n.epsilon = epsilon
        return self._add(n, in_node.shape)

    def create_leaky_relu(self, in_node):
        return self._add(_Node("leaky_relu", in_node), in_node.shape)

    def create_max_pool2d(self, in_node, ksize, strides, padding):
        b, h, w, c = in_node.shape
        n = _Node("max_pool2d", in_node)
        n.ksize, n.strides = ksize, strides
        n.oh, n.pt, n.pb = out_pads(h, ksize[1], strides[1], padding)
        n.ow, n.pl, n.pr = out_pads(w, ksize[2], strides[2], padding)
        return self._add(n, (b, n.oh, n.ow, c))

    def set_out_node(self, node):
        self.out_node = node

    # -- execution --------------------------------------------------------------------
    def run(self, x, keep_argmax=False):
        """Runs every node; returns the out node's result.  All results stay on the nodes."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert tuple(x.shape) == self.in_shape
        if self.backend != "oracle":
            assert x.shape[0] == 1, "the compiled reference is only correct for batch 1"
        for n in self.nodes:
            if n.kind == "input":
                n.result = x
            else:
                getattr(self, "_run_" + n.kind)(n, keep_argmax)
        return (self.out_node or self.nodes[-1]).result

    def _run_conv2d(self, n, _):
        src = n.in_node.result
        pad = np.pad(src, [(0, 0), (n.pt, n.pb), (n.pl, n.pr), (0, 0)], 'constant')
        b, ih, iw, ic = pad.shape
        kh, kw = n.kernel.shape[:2]
        od = n.shape[3]
        res = np.zeros(n.shape, dtype=np.float32)
        ci = ctypes.c_int
        if self.backend == "ref_avx":
            args = np.array([n.oh, n.ow, od, ih, iw, ic, kh, kw, n.strides[1], n.strides[2]], dtype=np.int32)
            self.lib.conv2d_pthread(fp(pad), fp(n.kernel), fp(res), ci(b), args.ctypes.data_as(IP))
        else:
            col = np.zeros((n.oh * n.ow, ic * kh * kw), dtype=np.float32) if self.backend == "oracle" \
                else np.zeros((b, n.oh * n.ow, ic * kh * kw), dtype=np.float32)
            f = self.lib.oracle_conv2d_mul if self.backend == "oracle" else self.lib.conv2d_mul
            f(fp(pad), fp(col), fp(n.kernel_r), fp(res), ci(b), ci(n.oh), ci(n.ow), ci(od),
              ci(ih), ci(iw), ci(ic), ci(kh), ci(kw), ci(n.strides[1]), ci(n.strides[2]))
        n.result = res

    def _run_bias_add(self, n, _):
        src = n.in_node.result
        res = np.zeros(n.shape, dtype=np.float32)
        f = {"oracle": "oracle_bias_add", "ref_openblas": "bias_add", "ref_avx": "bias_add_pthread"}[self.backend]
        getattr(self.lib, f)(fp(src), fp(n.biases), fp(res), *map(ctypes.c_int, n.shape))
        n.result = res

    def _run_batch_norm(self, n, _):
        src = n.in_node.result
        res = np.zeros(n.shape, dtype=np.float32)
        dims = list(map(ctypes.c_int, n.shape))
        if self.backend == "ref_avx":
            denom = np.sqrt(n.variance + n.epsilon)       # proj3/dnn_avx.py:301-303
            alpha = (n.gamma / denom).astype(np.float32)
            beta = (alpha * n.mean).astype(np.float32)
            self.lib.batch_norm(fp(src), fp(alpha), fp(beta), fp(res), *dims)
        else:
            var = n.variance.copy()
            f = self.lib.oracle_batch_norm if self.backend == "oracle" else self.lib.batch_norm
            f(fp(src), fp(n.mean), fp(var), fp(n.gamma), ctypes.c_float(n.epsilon), fp(res), *dims)
        n.result = res

    def _run_leaky_relu(self, n, _):
        src = n.in_node.result
        res = np.zeros(n.shape, dtype=np.float32)
        f = self.lib.oracle_leaky_relu if self.backend == "oracle" else self.lib.leaky_relu
        f(fp(src), fp(res), *map(ctypes.c_int, n.shape))
        n.result = res

    def _run_max_pool2d(self, n, keep_argmax):
        src = n.in_node.result
        pad = np.pad(src, [(0, 0), (n.pt, n.pb), (n.pl, n.pr), (0, 0)], 'constant', constant_values=F32_MIN)
        res = np.zeros(n.shape, dtype=np.float32)
        ci = ctypes.c_int
        b, ih, iw, ic = pad.shape
        kh, kw, sh, sw = n.ksize[1], n.ksize[2], n.strides[1], n.strides[2]
        if self.backend == "oracle":
            arg = np.zeros(n.shape, dtype=np.int32) if keep_argmax else None
            self.lib.oracle_max_pool2d(fp(pad), fp(res), arg.ctypes.data_as(IP) if keep_argmax else None,
                                       ci(b), ci(n.oh), ci(n.ow), ci(ic), ci(ih), ci(iw), ci(ic),
                                       ci(kh), ci(kw), ci(sh), ci(sw))
            n.argmax = arg
        elif self.backend == "ref_openblas":
            self.lib.max_pool2d(fp(pad), fp(res), ci(b), ci(n.oh), ci(n.ow), ci(ic), ci(ih), ci(iw), ci(ic),
                                ci(kh), ci(kw), ci(sh), ci(sw), ci(n.pt), ci(n.pb), ci(n.pl), ci(n.pr))
        else:
            args = np.array([n.oh, n.ow, ic, ih, iw, ic, kh, kw, sh, sw], dtype=np.int32)
            self.lib.max_pool2d_pthread(fp(pad), fp(res), ci(b), args.ctypes.data_as(IP))
        n.result = res


def pool_argmax(x, ksize, stride, padding='SAME'):
    """Raster index (di*kw+dj) of the first maximum in each pooling window of NHWC `x`:
    the oracle's definition of the max-pool index (proj3/dnn_openblas.c:7,209-222)."""
    c = Chain("oracle")
    i = c.create_input(x.shape)
    p = c.create_max_pool2d(i, [1, ksize, ksize, 1], [1, stride, stride, 1], padding)
    c.run(x, keep_argmax=True)
    return p.result, p.argmax


def yolo_forward(weights, images, backend="oracle", keep=False):
    """Runs the YOLOv2-tiny chain image by image.  Returns [B,13,13,125] (and, with keep=True,
    the list of per-node results of the LAST image)."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "..", "dnn-inference-engine_b200"))
    from synth import build_yolov2tiny
    outs = []
    chain = None
    for i in range(images.shape[0]):
        chain = Chain(backend)
        build_yolov2tiny(chain, (1,) + tuple(images.shape[1:]), weights)
        outs.append(chain.run(images[i:i + 1]))
    out = np.concatenate(outs, axis=0)
    if keep:
        return out, [n.result for n in chain.nodes]
    return out
