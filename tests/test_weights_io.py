"""CPU tests of the weight containers: the reference's pickle layout and darknet's."""
import pickle

import numpy as np
import pytest

import weights_io
from synth import CHANNELS, KSIZES, make_weights


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert sorted(x) == sorted(y)
        for k in x:
            np.testing.assert_array_equal(x[k], y[k])
            assert y[k].dtype == np.float32 and y[k].flags["C_CONTIGUOUS"]


def test_pickle_round_trip_and_python2_style_file(tmp_path):
    w = make_weights(3)
    p = tmp_path / "y2t_weights.pickle"
    weights_io.save_pickle(str(p), w)
    _same(w, weights_io.load_pickle(str(p)))
    with open(p, "wb") as h:                       # what a float64 / list-valued producer would leave behind
        pickle.dump([{k: np.asarray(v, dtype=np.float64) for k, v in layer.items()} for layer in w], h, protocol=2)
    _same(w, weights_io.load_pickle(str(p)))


def test_darknet_round_trip(tmp_path):
    w = make_weights(4)
    p = tmp_path / "yolov2-tiny-voc.weights"
    weights_io.save_darknet(str(p), w, seen=123)
    _same(w, weights_io.load_darknet(str(p), CHANNELS, KSIZES))
    # darknet stores filters as [O][I][H][W]: check one element against the HWIO original
    raw = np.fromfile(p, dtype="<f4", offset=20)
    oc = CHANNELS[1]
    first_filters = raw[4 * oc: 4 * oc + oc * 3 * 9].reshape(oc, 3, 3, 3)
    assert first_filters[5, 2, 1, 0] == w[0]["kernel"][1, 0, 2, 5]


def test_rejects_malformed_containers(tmp_path):
    w = make_weights(5)
    bad = [dict(x) for x in w]
    del bad[2]["gamma"]
    with pytest.raises(ValueError, match="batch norm"):
        weights_io._validate(bad)
    bad = [dict(x) for x in w]
    bad[3]["kernel"] = bad[3]["kernel"][:, :, :-1, :]
    with pytest.raises(ValueError, match="input channels"):
        weights_io._validate(bad)
    p = tmp_path / "short.weights"
    weights_io.save_darknet(str(p), w)
    with open(p, "r+b") as h:
        h.truncate(1000)
    with pytest.raises(ValueError, match="too short"):
        weights_io.load_darknet(str(p), CHANNELS, KSIZES)
