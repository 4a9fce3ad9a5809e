"""CPU tests: the oracle (oracle/oracle.c + oracle/chain.py) against the golden vectors that
oracle/make_golden.py produced by running the reference's own code, and the compiled
reference libraries (when present) against the same vectors."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, normalised_error
import chain
from synth import build_yolov2tiny, make_images, make_weights

OPS = np.load(os.path.join(GOLDEN, "ops_small.npz"))


def _one(backend, x, fn):
    c = chain.Chain(backend)
    i = c.create_input(x.shape)
    fn(c, i)
    return c.run(x)


@pytest.mark.parametrize("backend", ["oracle", "ref_openblas"])
def test_small_ops_match_reference(oracle_libs, backend):
    # ref_avx is left out on purpose: its vector loops overrun buffers whose channel count is
    # not a multiple of 8 -- it is only driven with the YOLO shapes below.
    if not chain.available(backend):
        pytest.skip(backend + " library not built here")
    o = OPS
    close = dict(rtol=2e-6, atol=2e-6)
    for name, pad in [("conv3_same", 'SAME'), ("conv3_valid", 'VALID')]:
        y = _one(backend, o[name + "_x"], lambda c, i: c.create_conv2d(i, o["conv3_k"], [1, 1, 1, 1], pad))
        np.testing.assert_allclose(y, o[name + "_y"], **close)
    y = _one(backend, o["conv1_same_x"], lambda c, i: c.create_conv2d(i, o["conv1_k"], [1, 1, 1, 1], 'SAME'))
    np.testing.assert_allclose(y, o["conv1_same_y"], **close)
    y = _one(backend, o["bias_x"], lambda c, i: c.create_bias_add(i, o["bias_b"]))
    np.testing.assert_array_equal(y, o["bias_y"])
    y = _one(backend, o["bn_x"], lambda c, i: c.create_batch_norm(i, o["bn_mean"], o["bn_var"], o["bn_gamma"], 1e-5))
    np.testing.assert_array_equal(y, o["bn_y"])
    y = _one(backend, o["leaky_x"], lambda c, i: c.create_leaky_relu(i))
    np.testing.assert_array_equal(y, o["leaky_y"])
    for name, s in [("pool_s2_even", 2), ("pool_s2_odd", 2), ("pool_s1_same", 1), ("pool_ties", 2)]:
        y = _one(backend, o[name + "_x"], lambda c, i: c.create_max_pool2d(i, [1, 2, 2, 1], [1, s, s, 1], 'SAME'))
        np.testing.assert_array_equal(y, o[name + "_y"])


def test_pool_argmax_is_first_max_in_raster_order(oracle_libs):
    x = OPS["pool_ties_x"]
    y, arg = chain.pool_argmax(x, 2, 2)
    np.testing.assert_array_equal(y, OPS["pool_ties_y"])
    win = x.reshape(1, 3, 2, 3, 2, 4).transpose(0, 1, 3, 5, 2, 4).reshape(1, 3, 3, 4, 4)
    np.testing.assert_array_equal(arg, win.argmax(axis=-1))   # np.argmax = first maximum
    assert (np.take_along_axis(win, arg[..., None], -1)[..., 0] == y).all()


def test_mini_net_every_node(oracle_libs):
    g = np.load(os.path.join(GOLDEN, "mini_net.npz"))
    weights = make_weights(5, channels=[3, 8, 16, 10], ksizes=[3, 3, 1], negative_gamma=True)
    x = make_images(1, seed=55, hw=13)
    c = chain.Chain("oracle")
    cur = c.create_input(x.shape)
    for i, w in enumerate(weights):
        cur = c.create_conv2d(cur, w["kernel"], [1, 1, 1, 1], 'SAME')
        cur = c.create_bias_add(cur, w["biases"])
        if i < 2:
            cur = c.create_batch_norm(cur, w["moving_mean"], w["moving_variance"], w["gamma"], 1e-5)
            cur = c.create_leaky_relu(cur)
            cur = c.create_max_pool2d(cur, [1, 2, 2, 1], [1, 2 - i, 2 - i, 1], 'SAME')
    c.run(x)
    for i, n in enumerate(c.nodes[1:]):
        np.testing.assert_allclose(n.result, g["node_%02d" % i], rtol=1e-5, atol=1e-5, err_msg="node %d" % i)


@pytest.mark.parametrize("wseed,iseed", [(0, 100), (1, 101)])
@pytest.mark.parametrize("backend", ["oracle", "ref_openblas", "ref_avx"])
def test_yolo_end_to_end_matches_reference(oracle_libs, backend, wseed, iseed):
    if not chain.available(backend):
        pytest.skip(backend + " library not built here")
    g = np.load(os.path.join(GOLDEN, "yolo_w%d_i%d.npz" % (wseed, iseed)))
    out, nodes = chain.yolo_forward(make_weights(wseed), make_images(1, seed=iseed), backend, keep=True)
    assert out.shape == (1, 13, 13, 125)
    assert normalised_error(out, g["out_openblas"]) < 2e-5
    assert normalised_error(out, g["out_avx"]) < 2e-5
    # layer-wise, the reference's own verification convention (layer_N tensors)
    assert len(nodes) - 1 == len(g["node_sum"]) == 40
    for i, r in enumerate(nodes[1:]):
        assert tuple(r.shape) == tuple(g["node_shape"][i])
        s = r.ravel()[:: max(1, r.size // 64)][:64]
        scale = max(float(g["node_absmax"][i]), 1e-6)
        assert np.abs(s - g["node_sample"][i]).max() / scale < 2e-5, "node %d" % i
        assert abs(float(np.abs(r).max()) - float(g["node_absmax"][i])) / scale < 2e-5
    if backend == "ref_openblas":
        np.testing.assert_array_equal(out, g["out_openblas"])   # same code, same box class: bit-equal


def test_oracle_handles_batches_the_reference_cannot(oracle_libs):
    """The reference is wrong for batch > 1 (App. B2/B3); the restatement is per-image exact."""
    w = make_weights(3, channels=[3, 8, 12], ksizes=[3, 1])
    x = make_images(3, seed=9, hw=12)
    c = chain.Chain("oracle"); build_yolov2tiny(c, x.shape, w)
    yb = c.run(x)
    for i in range(3):
        c1 = chain.Chain("oracle"); build_yolov2tiny(c1, (1,) + x.shape[1:], w)
        np.testing.assert_array_equal(c1.run(x[i:i + 1])[0], yb[i])
