"""dnn_b200 -- drop-in for proj3/dnn_openblas.py (the module proj3/yolov2tiny.py:5 imports):

    from dnn_b200 import DnnGraphBuilder, DnnInferenceEngine

Same builder methods, node classes, constructor checks (ValueError) and return types as the
reference host layer (proj3/dnn_openblas.py:23-299); underneath, one C-ABI library built for
sm_100a (libdnn_b200.so, include/dnn_b200.h) instead of libdnn_openblas.so.

What differs from the reference, on purpose:
  * DnnInferenceEngine.run() recognises the chain Conv2D -> BiasAdd -> [BatchNorm -> LeakyReLU]
    -> [MaxPool2D] and runs each such run of nodes as ONE fused tcgen05 kernel on a network whose
    weights and activations stay in HBM; only the input and the out-node result cross PCIe.
    The reference runs 41 nodes one by one with a numpy round trip (and an np.save) per node
    (proj3/dnn_openblas.py:29-57).
  * debug=True keeps the reference's layer-wise workflow: every node runs on its own through the
    op-level entry points (fp32 CUDA kernels) and its result is saved to
    ./intermediate/layer_N.npy (0-based, as proj3/dnn_avx.py:54-58 does).
  * any batch size works (the reference is only correct for batch 1).
  * node construction and run() do not print; a node's .result is allocated on first access.
  * extra keyword arguments: precision ('tf32' | 'bf16' | 'fp32') and device.
There is no CPU path: importing this module fails if libdnn_b200.so is missing, and every run
fails if no CUDA device is present.
"""
import math
import os

import networkx as nx
import numpy as np

import _lib
from _lib import check, f32c, fptr, lib

F32_MIN = np.finfo(np.float32).min


class DnnInferenceEngine(object):
    def __init__(self, graph, debug=False, precision="tf32", device=0, latency=None, devices=None, latency_widths=None):
        """devices=[0, 1, ...]: batch-sharded over several GPUs (sharded.MultiGpuEngine: contiguous slices of the
        batch, weights replicated once per device, no collective; the result is assembled in one host array)."""
        self.g = graph
        self.debug = debug
        self.precision = precision
        self.device = device if devices is None else list(devices)[0]
        self.devices = None if devices is None else list(devices)
        self.latency = latency       # True: plan for single-image latency (see _lib.Net); None: DNNB_LATENCY_MODE decides
        self.latency_widths = latency_widths   # per-block N-tile widths of that plan; latency='tune' measures them (see _build_net)
        self.save_dir = os.path.join(os.getcwd(), "intermediate")
        self._net = None
        self._net_key = None
        self.fused_blocks = None     # list of node lists, once fused

    # -- execution order: the reference's BFS with predecessor check (proj3/dnn_openblas.py:29-57)
    def _ordered_nodes(self):
        order, done = [], set()
        currents = [self.g.in_node]
        while currents:
            nexts = []
            for cur in currents:
                if cur in done:
                    continue
                waiting = [p for p in self.g.G.predecessors(cur) if p not in done]
                if waiting:
                    nexts.extend(waiting)
                    continue
                order.append(cur)
                done.add(cur)
                nexts.extend(self.g.G.successors(cur))
            currents = nexts
        return order

    def _fuse(self):
        """Splits a linear chain into conv blocks; returns None when the graph is anything else."""
        order = self._ordered_nodes()
        if not order or not isinstance(order[0], Input) or self.g.out_node is None:
            return None
        for a, b in zip(order, order[1:]):
            if b.in_node is not a or self.g.G.out_degree(a) != 1:
                return None
        if order[-1] is not self.g.out_node:
            return None
        blocks, i, nodes = [], 1, order
        while i < len(nodes):
            if not isinstance(nodes[i], Conv2D) or list(nodes[i].strides[1:3]) != [1, 1]:
                return None
            blk = {"nodes": [nodes[i]], "conv": nodes[i]}
            i += 1
            for kind, key in ((BiasAdd, "bias"), (BatchNorm, "bn"), (LeakyReLU, "leaky"), (MaxPool2D, "pool")):
                if i < len(nodes) and isinstance(nodes[i], kind):
                    blk[key] = nodes[i]
                    blk["nodes"].append(nodes[i])
                    i += 1
            p = blk.get("pool")
            if p is not None and (p.ksize[1] != p.ksize[2] or p.strides[1] != p.strides[2]):
                return None
            blocks.append(blk)
        return blocks

    def _build_net(self, blocks, device=None, batch=None):
        spec = []
        for blk in blocks:
            conv, bn, pool = blk["conv"], blk.get("bn"), blk.get("pool")
            spec.append({
                "kernel": conv.kernel, "padding": conv.padding,
                "biases": blk["bias"].biases if "bias" in blk else None,
                "mean": bn.mean if bn else None, "variance": bn.variance if bn else None,
                "gamma": bn.gamma if bn else None, "epsilon": bn.epsilon if bn else 0.0,
                "leaky": "leaky" in blk,
                "pool": None if pool is None else (pool.ksize[1], pool.strides[1], pool.padding),
            })
        in_shape = list(self.g.in_node.in_shape)
        if batch is not None:
            in_shape[0] = batch
        dev = self.device if device is None else device
        if self.latency == "tune":
            # measured plan: the N-tile widths of the latency plan are timed on this device once (a few seconds) and
            # kept in self.latency_widths; pass them back as latency_widths= to pin the plan without measuring again
            if self.latency_widths is None:
                self.latency_widths = _lib.tune_latency_widths(in_shape, spec, self.precision, dev)["widths"]
            return _lib.Net(in_shape, spec, precision=self.precision, device=dev, latency=True, latency_widths=self.latency_widths)
        return _lib.Net(in_shape, spec, precision=self.precision, device=dev, latency=self.latency,
                        latency_widths=self.latency_widths)

    def _build_sharded(self, blocks):
        import sharded
        return sharded.MultiGpuEngine(lambda dev, n: self._build_net(blocks, dev, n), self.g.in_node.in_shape[0], self.devices)

    def run(self, tin):
        self.g.in_node.set_input(tin)
        if not self.debug:
            blocks = self._fuse()
            if blocks is not None:
                if self._net is None:
                    self._net = self._build_sharded(blocks) if self.devices else self._build_net(blocks)
                    self.fused_blocks = [b["nodes"] for b in blocks]
                out = self._net.run(f32c(tin)) if self.devices else self._net.forward(tin)
                self.g.out_node.result = out
                return out
        # node-by-node (debug mode, or a graph that is not a chain of conv blocks)
        out = {}
        counter = 0
        for node in self._ordered_nodes():
            node.run()
            if not isinstance(node, Input):
                if self.debug:
                    os.makedirs(self.save_dir, exist_ok=True)
                    np.save(os.path.join(self.save_dir, "layer_{}.npy".format(counter)), node.result)
                counter += 1
            if self.g.is_out_node(node):
                out = node.result
        return out

    def block_results(self):
        """After a fused run: {last node of each block: its result}, for layer-wise comparison at
        block granularity (the tensors inside a fused block never exist)."""
        if self._net is None:
            raise RuntimeError("no fused run yet")
        if self.devices:
            raise RuntimeError("block results are kept per device: run on one device to compare layers")
        return {nodes[-1]: self._net.block_output(i) for i, nodes in enumerate(self.fused_blocks)}

    def kernel_plan(self):
        if self._net is None:
            blocks = self._fuse()
            self._net = self._build_sharded(blocks) if self.devices else self._build_net(blocks)
            self.fused_blocks = [b["nodes"] for b in blocks]
        net = next(n for n in self._net.nets if n is not None) if self.devices else self._net
        return [net.block_info(i) for i in range(net.num_blocks)]

    def close(self):
        if self._net is not None:
            self._net.close()
            self._net = None


class DnnGraphBuilder(object):
    def __init__(self):
        self.G = nx.DiGraph()
        self.name_num = {"conv2d": 0, "bias_add": 0, "max_pool2d": 0, "batch_norm": 0, "leaky_relu": 0, "input": 0}
        self.in_node = None
        self.out_node = None

    def set_in_node(self, node):
        self.in_node = node

    def set_out_node(self, node):
        self.out_node = node

    def is_out_node(self, node):
        return self.out_node is node

    def get_name(self, layer_name):
        name = layer_name + "_" + str(self.name_num[layer_name])
        self.name_num[layer_name] += 1
        return name

    def _link(self, in_node, out_node):
        self.G.add_edge(in_node, out_node)
        return out_node

    def create_conv2d(self, in_node, kernel, strides, padding):
        return self._link(in_node, Conv2D(self.get_name("conv2d"), in_node, kernel, strides, padding))

    def create_bias_add(self, in_node, biases):
        return self._link(in_node, BiasAdd(self.get_name("bias_add"), in_node, biases))

    def create_max_pool2d(self, in_node, ksize, strides, padding):
        return self._link(in_node, MaxPool2D(self.get_name("max_pool2d"), in_node, ksize, strides, padding))

    def create_batch_norm(self, in_node, mean, variance, gamma, epsilon):
        return self._link(in_node, BatchNorm(self.get_name("batch_norm"), in_node, mean, variance, gamma, epsilon))

    def create_leaky_relu(self, in_node):
        return self._link(in_node, LeakyReLU(self.get_name("leaky_relu"), in_node))

    def create_input(self, in_shape):
        out_node = Input(self.get_name("input"), in_shape)
        self.G.add_node(out_node)
        self.set_in_node(out_node)  # one input, as in the reference
        return out_node


class DnnNode(object):
    """Base: `shape` is known at construction; `result` is a float32 NHWC host array, allocated
    (zeros, as the reference does at build time) the first time somebody looks at it."""
    shape = None
    _result = None

    @property
    def result(self):
        if self._result is None:
            self._result = np.zeros(self.shape, dtype=np.float32)
        return self._result

    @result.setter
    def result(self, value):
        self._result = value

    def run(self):
        self.result = None


def get_out_pads(in_size, filter_size, stride_size, padding):
    """TensorFlow SAME / VALID output size and padding (proj3/dnn_openblas.py:127-142)."""
    assert padding == 'SAME' or padding == 'VALID'
    if padding == 'SAME':
        out_size = math.ceil(float(in_size) / float(stride_size))
        pad_size = max((out_size - 1) * stride_size + filter_size - in_size, 0)
        pad_front = pad_size // 2
        pad_back = pad_size - pad_front
    else:
        out_size = math.ceil(float(in_size - filter_size + 1) / float(stride_size))
        pad_front = 0
        pad_back = 0
    return out_size, pad_front, pad_back


class Conv2D(DnnNode):
    def __init__(self, name, in_node, kernel, strides, padding):
        batch, in_h, in_w, ic = in_node.shape
        kh, kw, kernel_ic, od = kernel.shape
        if kernel_ic != ic:
            raise ValueError
        if not (padding == 'SAME' or padding == 'VALID'):
            raise ValueError
        oh, self.pad_top, self.pad_bottom = get_out_pads(in_h, kh, strides[1], padding)
        ow, self.pad_left, self.pad_right = get_out_pads(in_w, kw, strides[2], padding)
        self.in_node = in_node
        self.kernel = f32c(kernel)
        self.strides = strides
        self.padding = padding
        self.shape = (batch, oh, ow, od)
        self.name = name

    def run(self):
        # the reference's own sequence (proj3/dnn_openblas.py:174-188): zero-pad on the host, then
        # conv2d_mul(padded input, scratch, kernel_r[K, od], result, shapes...)
        x = np.pad(f32c(self.in_node.result),
                   [(0, 0), (self.pad_top, self.pad_bottom), (self.pad_left, self.pad_right), (0, 0)], 'constant')
        kh, kw, ic, od = self.kernel.shape
        kernel_r = np.ascontiguousarray(self.kernel.transpose(2, 0, 1, 3).reshape(-1, od))
        res = np.zeros(self.shape, dtype=np.float32)
        check(lib.conv2d_mul(fptr(x), None, fptr(kernel_r), fptr(res), *res.shape, *x.shape[1:], kh, kw,
                             self.strides[1], self.strides[2]))
        self.result = res


class BiasAdd(DnnNode):
    def __init__(self, name, in_node, biases):
        if not (biases.ndim == 1 and in_node.shape[-1] == biases.shape[0]):
            raise ValueError
        self.in_node = in_node
        self.biases = f32c(biases)
        self.shape = tuple(in_node.shape)
        self.name = name

    def run(self):
        res = np.zeros(self.shape, dtype=np.float32)
        check(lib.bias_add(fptr(f32c(self.in_node.result)), fptr(self.biases), fptr(res), *res.shape))
        self.result = res


class MaxPool2D(DnnNode):
    def __init__(self, name, in_node, ksize, strides, padding):
        if not (padding == 'SAME' or padding == 'VALID'):
            raise ValueError
        batch, in_h, in_w, ch = in_node.shape
        oh, self.pad_top, self.pad_bottom = get_out_pads(in_h, ksize[1], strides[1], padding)
        ow, self.pad_left, self.pad_right = get_out_pads(in_w, ksize[2], strides[2], padding)
        self.in_node = in_node
        self.ksize = ksize
        self.strides = strides
        self.padding = padding
        self.shape = (batch, oh, ow, ch)
        self.name = name

    def run(self):
        # proj3/dnn_openblas.py:230-242: pad with finfo(float32).min on the host, then max_pool2d
        x = np.pad(f32c(self.in_node.result),
                   [(0, 0), (self.pad_top, self.pad_bottom), (self.pad_left, self.pad_right), (0, 0)],
                   'constant', constant_values=F32_MIN)
        res = np.zeros(self.shape, dtype=np.float32)
        check(lib.max_pool2d(fptr(x), fptr(res), *res.shape, *x.shape[1:], self.ksize[1], self.ksize[2],
                             self.strides[1], self.strides[2]))
        self.result = res


class BatchNorm(DnnNode):
    def __init__(self, name, in_node, mean, variance, gamma, epsilon):
        if not all(arg.ndim == 1 and in_node.shape[-1] == arg.shape[0] for arg in [mean, variance, gamma]):
            raise ValueError
        self.in_node = in_node
        self.mean = f32c(mean)
        self.variance = f32c(variance)
        self.gamma = f32c(gamma)
        self.epsilon = epsilon
        self.shape = tuple(in_node.shape)
        self.name = name

    def run(self):
        res = np.zeros(self.shape, dtype=np.float32)
        check(lib.batch_norm(fptr(f32c(self.in_node.result)), fptr(self.mean), fptr(self.variance), fptr(self.gamma),
                             float(self.epsilon), fptr(res), *res.shape))
        self.result = res


class LeakyReLU(DnnNode):
    def __init__(self, name, in_node):
        self.in_node = in_node
        self.shape = tuple(in_node.shape)
        self.name = name

    def run(self):
        res = np.zeros(self.shape, dtype=np.float32)
        check(lib.leaky_relu(fptr(f32c(self.in_node.result)), fptr(res), *res.shape))
        self.result = res


class Input(DnnNode):
    def __init__(self, name, in_shape):
        self.name = name
        self.in_shape = in_shape
        self.shape = tuple(in_shape)

    def set_input(self, tensor):
        assert tuple(self.in_shape) == tuple(tensor.shape)
        self.result = tensor

    def run(self):
        pass
