"""The import swap INTEGRATION.md section A documents, executed: the reference's own YOLO_V2_TINY class
(proj3/yolov2tiny.py:7-83, imported from /root/reference, never copied) built on top of dnn_b200 in place of
dnn_openblas.  Runs where /root/reference exists (the build container); graph construction needs no GPU."""
import importlib
import os
import sys

import numpy as np
import pytest

import dnn_b200
import weights_io
from synth import build_yolov2tiny, make_weights

REF = "/root/reference/cs492-projects/proj3"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")


@pytest.fixture()
def ref_yolov2tiny(monkeypatch):
    """`from dnn_openblas import DnnGraphBuilder, DnnInferenceEngine` (yolov2tiny.py:5) resolves to dnn_b200."""
    monkeypatch.setitem(sys.modules, "dnn_openblas", dnn_b200)
    monkeypatch.syspath_prepend(REF)
    sys.modules.pop("yolov2tiny", None)
    mod = importlib.import_module("yolov2tiny")
    yield mod
    sys.modules.pop("yolov2tiny", None)


def test_reference_model_class_builds_its_graph_on_dnn_b200(ref_yolov2tiny, tmp_path):
    weights = make_weights(0)
    path = tmp_path / "y2t_weights.pickle"
    weights_io.save_pickle(str(path), weights)

    class Model(ref_yolov2tiny.YOLO_V2_TINY):
        def get_y2t_w(self):                      # yolov2tiny.py:17-22 refuses Python != 2.7 / 3.5
            return weights_io.load_pickle(self.weight_pickle)

    m = Model([1, 416, 416, 3], str(path), False)
    assert isinstance(m.g, dnn_b200.DnnGraphBuilder) and isinstance(m.sess, dnn_b200.DnnInferenceEngine)
    # same graph as the one the tests and the bench build with synth.build_yolov2tiny
    g = dnn_b200.DnnGraphBuilder()
    mine = build_yolov2tiny(g, [1, 416, 416, 3], weights)
    theirs = m.sess._ordered_nodes()
    assert [n.name for n in theirs] == [n.name for n in mine]
    assert [tuple(n.shape) for n in theirs] == [tuple(n.shape) for n in mine]
    assert theirs[-1] is m.g.out_node and theirs[-1].name == "bias_add_8" and tuple(theirs[-1].shape) == (1, 13, 13, 125)
    for a, b in zip(theirs, mine):
        for attr in ("kernel", "biases", "mean", "variance", "gamma"):
            if hasattr(b, attr):
                np.testing.assert_array_equal(getattr(a, attr), getattr(b, attr))
    # and the engine sees the nine fusable conv blocks of proj3/yolov2tiny.py:30-77
    blocks = m.sess._fuse()
    assert blocks is not None and len(blocks) == 9
    assert [len(b["nodes"]) for b in blocks] == [5, 5, 5, 5, 5, 5, 4, 4, 2]
    pool5 = blocks[5]["pool"]
    assert pool5.strides == [1, 1, 1, 1] and (pool5.pad_top, pool5.pad_bottom) == (0, 1)


def test_reference_postprocessing_is_importable_next_to_the_swap(ref_yolov2tiny):
    # the module's postprocessing (yolov2tiny.py:94-177) is untouched by the swap and still the oracle's anchor
    quiet = np.full((13, 13, 125), -30.0, np.float32)
    assert ref_yolov2tiny.postprocessing(quiet) == []
