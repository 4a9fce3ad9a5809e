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
    weights_io.save_darknet(str(p), w, seen=123, convert=False)           # vectors written and read as they are
    _same(w, weights_io.load_darknet(str(p), CHANNELS, KSIZES, convert=False))
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


# ------------------------------------------------------------------------------------------------
# Byte-level fixtures that weights_io's own writers did not produce (tests/golden/make_weight_fixtures.py
# assembles them with struct.pack and a hand-written protocol-2 opcode stream).
import importlib.util
import os

from conftest import GOLDEN

_spec = importlib.util.spec_from_file_location("make_weight_fixtures", os.path.join(GOLDEN, "make_weight_fixtures.py"))
fixtures = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(fixtures)


def _darknet_layer0(conv):
    """darknet's own arithmetic for the batch-normalised layer (normalize_cpu, scale_bias, add_bias), float64."""
    l0, _ = fixtures.layer_values()
    s, m, v, b = (np.array(l0[k], dtype=np.float64) for k in ("scales", "mean", "variance", "biases"))
    return s * (conv - m) / (np.sqrt(v) + 1e-6) + b


def _reference_layer0(conv, w):
    """the reference graph's arithmetic (bias_add then batch_norm, proj3/dnn_openblas.c:9-65), float64."""
    g, m, v, b = (w[k].astype(np.float64) for k in ("gamma", "moving_mean", "moving_variance", "biases"))
    return (conv + b - m) / np.sqrt(v + 1e-5) * g


@pytest.mark.parametrize("name,seen_bytes", [("tiny_darknet_v01.weights", 4), ("tiny_darknet_v02.weights", 8)])
def test_darknet_fixture_both_header_versions_and_bn_convention(name, seen_bytes):
    path = os.path.join(GOLDEN, name)
    assert os.path.getsize(path) == 12 + seen_bytes + 4 * (4 * 4 + 108 + 5 + 20)
    assert open(path, "rb").read() == fixtures.darknet_bytes(0, 1 if seen_bytes == 4 else 2, 77 if seen_bytes == 4 else 1234567890123)
    w = weights_io.load_darknet(path, fixtures.CHANNELS, fixtures.KSIZES)
    l0, l1 = fixtures.layer_values()
    # filters [O][I][H][W] -> HWIO, element by element against the generating formula
    for o, i, h, x in [(0, 0, 0, 0), (3, 2, 1, 0), (1, 1, 2, 2), (2, 0, 1, 1)]:
        assert w[0]["kernel"][h, x, i, o] == (o * 27 + i * 9 + h * 3 + x - 50) / 8.0
    assert w[1]["kernel"].shape == (1, 1, 4, 5) and w[1]["kernel"][0, 0, 3, 4] == (4 * 4 + 3 - 9) / 4.0
    np.testing.assert_array_equal(w[1]["biases"], np.array(l1["biases"], np.float32))
    assert "gamma" not in w[1]
    # the converted layer computes what darknet computes, for any conv value -- including the channel whose
    # scale is negative and the one whose scale is zero (a constant)
    conv = np.random.default_rng(0).standard_normal((64, 4)) * 3
    want, got = _darknet_layer0(conv), _reference_layer0(conv, w[0])
    assert np.abs(got - want).max() < 2e-6 * np.abs(want).max()
    # ... which copying the vectors across (convert=False, what the importer used to do) does not
    raw = weights_io.load_darknet(path, fixtures.CHANNELS, fixtures.KSIZES, convert=False)
    np.testing.assert_array_equal(raw[0]["gamma"], np.array(l0["scales"], np.float32))
    assert np.abs(_reference_layer0(conv, raw[0]) - want).max() > 0.1


def test_darknet_writer_inverts_the_conversion(tmp_path):
    w = make_weights(6, negative_gamma=True)
    for version, seen_bytes in (((0, 2, 0), 8), ((0, 1, 0), 4)):
        p = tmp_path / ("v%d.weights" % version[1])
        weights_io.save_darknet(str(p), w, seen=5, version=version)
        back = weights_io.load_darknet(str(p), CHANNELS, KSIZES)
        for a, b in zip(w, back):
            np.testing.assert_array_equal(a["kernel"], b["kernel"])
            for k in a:
                if k != "kernel":       # through float32 files: the affine map is preserved, not every bit
                    np.testing.assert_allclose(b[k], a[k], rtol=2e-6, atol=1e-7)
        # on disk the vectors are darknet's: bias after the scale
        raw = weights_io.load_darknet(str(p), CHANNELS, KSIZES, convert=False)
        sd = np.sqrt(w[0]["moving_variance"].astype(np.float64) + 1e-5)
        np.testing.assert_allclose(raw[0]["biases"], w[0]["biases"] * w[0]["gamma"] / sd, rtol=1e-6)


def test_python2_protocol2_pickle_fixture():
    path = os.path.join(GOLDEN, "tiny_py2_protocol2.pickle")
    blob = open(path, "rb").read()
    assert blob == fixtures.py2_pickle_bytes() and blob[:2] == b"\x80\x02"
    with pytest.raises(UnicodeDecodeError):            # a Python-2 pickle: the default ASCII decode fails, as in
        pickle.loads(blob)                             # proj3/yolov2tiny.py:19-22, which is why it passes latin1
    w = weights_io.load_pickle(path)
    l0, l1 = fixtures.layer_values()
    assert [sorted(x) for x in w] == [sorted(weights_io.KEYS_BN), ["biases", "kernel"]]
    assert w[0]["kernel"].shape == (3, 3, 3, 4) and w[0]["kernel"][1, 0, 2, 3] == 6.5
    np.testing.assert_array_equal(w[0]["gamma"], np.array(l0["scales"], np.float32))       # reference convention: as stored
    np.testing.assert_array_equal(w[0]["moving_variance"], np.array(l0["variance"], np.float32))
    np.testing.assert_array_equal(w[1]["biases"], np.array(l1["biases"], np.float32))
    assert all(v.dtype == np.float32 and v.flags["C_CONTIGUOUS"] for x in w for v in x.values())
