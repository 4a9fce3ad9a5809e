"""CPU tests of the host layer (dnn_b200.py) and of the C-ABI library's surface: names,
signatures and error behaviour mirror proj3/dnn_openblas.py; no compute runs without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
import dnn_b200
import _lib
from synth import CHANNELS, build_yolov2tiny, make_weights


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "dnn_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = re.findall(r"\b(?:int|const char \*)\s*\*?\s*([a-z_0-9]+)\s*\(", header)
    declared = [d for d in declared if d not in ("defined",)]
    assert len(declared) >= 25
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert sorted(declared) == sorted(_lib.EXPORTS)
    assert _lib.lib.dnnb_version() == 100


def test_reference_symbols_are_exported_for_a_cdll_swap():
    # proj3/dnn_openblas.py binds these five names from whatever .so it loads
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in ("conv2d_mul", "bias_add", "batch_norm", "max_pool2d", "leaky_relu"):
        assert hasattr(raw, name)


def test_get_out_pads_matches_tf_rules():
    assert dnn_b200.get_out_pads(416, 3, 1, 'SAME') == (416, 1, 1)
    assert dnn_b200.get_out_pads(13, 1, 1, 'SAME') == (13, 0, 0)
    assert dnn_b200.get_out_pads(416, 2, 2, 'SAME') == (208, 0, 0)
    assert dnn_b200.get_out_pads(13, 2, 1, 'SAME') == (13, 0, 1)      # the stride-1 pool of layer 5
    assert dnn_b200.get_out_pads(13, 2, 2, 'SAME') == (7, 0, 1)
    assert dnn_b200.get_out_pads(13, 3, 1, 'VALID') == (11, 0, 0)
    with pytest.raises(AssertionError):
        dnn_b200.get_out_pads(13, 3, 1, 'FULL')


def test_builder_api_names_shapes_and_graph():
    g = dnn_b200.DnnGraphBuilder()
    nodes = build_yolov2tiny(g, [2, 416, 416, 3], make_weights(0))
    assert len(nodes) == 41 and g.G.number_of_nodes() == 41 and g.G.number_of_edges() == 40
    assert g.in_node is nodes[0] and g.is_out_node(nodes[-1]) and not g.is_out_node(nodes[0])
    assert nodes[1].name == "conv2d_0" and nodes[-1].name == "bias_add_8" and nodes[5].name == "max_pool2d_0"
    assert nodes[-1].shape == (2, 13, 13, 125)
    assert nodes[5].shape == (2, 208, 208, 16)
    pool5 = [n for n in nodes if n.name == "max_pool2d_5"][0]
    assert pool5.shape == (2, 13, 13, 512) and (pool5.pad_top, pool5.pad_bottom) == (0, 1)
    # .result exists before any run, zeros of the node's shape, like the reference's build-time buffers
    r = nodes[-1].result
    assert r.shape == (2, 13, 13, 125) and r.dtype == np.float32 and not r.any()


def test_constructor_checks_raise_value_error():
    g = dnn_b200.DnnGraphBuilder()
    x = g.create_input([1, 8, 8, 4])
    k = np.zeros((3, 3, 5, 6), np.float32)
    with pytest.raises(ValueError):
        g.create_conv2d(x, k, [1, 1, 1, 1], 'SAME')                  # channel mismatch
    with pytest.raises(ValueError):
        g.create_conv2d(x, np.zeros((3, 3, 4, 6), np.float32), [1, 1, 1, 1], 'FULL')
    with pytest.raises(ValueError):
        g.create_bias_add(x, np.zeros(5, np.float32))
    with pytest.raises(ValueError):
        g.create_bias_add(x, np.zeros((4, 1), np.float32))
    with pytest.raises(ValueError):
        g.create_max_pool2d(x, [1, 2, 2, 1], [1, 2, 2, 1], 'full')
    with pytest.raises(ValueError):
        g.create_batch_norm(x, np.zeros(4, np.float32), np.zeros(3, np.float32), np.zeros(4, np.float32), 1e-5)
    inp = dnn_b200.Input("input_9", [1, 2, 2, 1])
    with pytest.raises(AssertionError):
        inp.set_input(np.zeros((1, 2, 2, 2), np.float32))


def test_fusion_pass_groups_the_yolo_chain_into_nine_blocks():
    g = dnn_b200.DnnGraphBuilder()
    build_yolov2tiny(g, [1, 416, 416, 3], make_weights(0))
    eng = dnn_b200.DnnInferenceEngine(g, False)
    blocks = eng._fuse()
    assert [len(b["nodes"]) for b in blocks] == [5, 5, 5, 5, 5, 5, 4, 4, 2]
    assert all("bias" in b for b in blocks)
    assert [("pool" in b) for b in blocks] == [True] * 6 + [False] * 3
    assert blocks[5]["pool"].strides == [1, 1, 1, 1]
    assert [b["conv"].kernel.shape[3] for b in blocks] == CHANNELS[1:]
    order = eng._ordered_nodes()
    assert len(order) == 41 and order[0] is g.in_node and order[-1] is g.out_node


def test_non_chain_graph_is_not_fused():
    g = dnn_b200.DnnGraphBuilder()
    x = g.create_input([1, 8, 8, 4])
    a = g.create_leaky_relu(x)          # does not start with a conv
    g.set_out_node(a)
    assert dnn_b200.DnnInferenceEngine(g, False)._fuse() is None


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="libdnn_b200"):
        _lib.device_count()
    x = np.zeros((1, 2, 2, 1), np.float32)
    with pytest.raises(RuntimeError, match="libdnn_b200"):
        _lib.check(_lib.lib.leaky_relu(_lib.fptr(x), _lib.fptr(x.copy()), 1, 2, 2, 1))
    g = dnn_b200.DnnGraphBuilder()
    build_yolov2tiny(g, [1, 32, 32, 3], make_weights(0, channels=[3, 8, 4], ksizes=[3, 1]))
    with pytest.raises(RuntimeError, match="libdnn_b200"):
        dnn_b200.DnnInferenceEngine(g, False).run(np.zeros((1, 32, 32, 3), np.float32))
