"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(libdnn_b200.so via ctypes) and is checked against the CPU oracle (oracle/) on the same seeded
inputs, and against the committed golden vectors the reference itself produced.

Tolerances (BASELINE.json north_star): normalised error max|out-ref|/max|ref| <= 1e-2 for TF32,
<= 3e-2 for BF16 on the [13,13,125] output; op-level fp32 entry points are bit-exact except the
convolution (summation order), which is held to 1e-5.
"""
import ctypes
import os

import numpy as np
import pytest

from conftest import GOLDEN, normalised_error
import _lib
import chain
import dnn_b200
from synth import CHANNELS, KSIZES, POOLS, build_yolov2tiny, make_images, make_weights

pytestmark = pytest.mark.gpu

TOL = {"tf32": 1e-2, "bf16": 3e-2, "fp32": 1e-4}
OPS = np.load(os.path.join(GOLDEN, "ops_small.npz"))


def _single(x, fn, debug=True):
    g = dnn_b200.DnnGraphBuilder()
    i = g.create_input(list(x.shape))
    g.set_out_node(fn(g, i))
    return dnn_b200.DnnInferenceEngine(g, debug).run(x)


# ------------------------------------------------------------------------------- op level
def test_op_level_entry_points_against_reference_goldens(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    o = OPS
    for name, pad in [("conv3_same", 'SAME'), ("conv3_valid", 'VALID')]:
        y = _single(o[name + "_x"], lambda g, i: g.create_conv2d(i, o["conv3_k"], [1, 1, 1, 1], pad))
        np.testing.assert_allclose(y, o[name + "_y"], rtol=1e-5, atol=1e-5)
    y = _single(o["conv1_same_x"], lambda g, i: g.create_conv2d(i, o["conv1_k"], [1, 1, 1, 1], 'SAME'))
    np.testing.assert_allclose(y, o["conv1_same_y"], rtol=1e-5, atol=1e-5)
    y = _single(o["bias_x"], lambda g, i: g.create_bias_add(i, o["bias_b"]))
    np.testing.assert_array_equal(y, o["bias_y"])
    y = _single(o["bn_x"], lambda g, i: g.create_batch_norm(i, o["bn_mean"], o["bn_var"], o["bn_gamma"], 1e-5))
    np.testing.assert_array_equal(y, o["bn_y"])
    y = _single(o["leaky_x"], lambda g, i: g.create_leaky_relu(i))
    np.testing.assert_array_equal(y, o["leaky_y"])
    for name, s in [("pool_s2_even", 2), ("pool_s2_odd", 2), ("pool_s1_same", 1), ("pool_ties", 2)]:
        y = _single(o[name + "_x"], lambda g, i: g.create_max_pool2d(i, [1, 2, 2, 1], [1, s, s, 1], 'SAME'))
        np.testing.assert_array_equal(y, o[name + "_y"])
    assert sorted(os.listdir(tmp_path / "intermediate"))[0] == "layer_0.npy"   # debug dump convention


def test_op_level_bit_exact_on_larger_random_tensors(oracle_libs):
    rng = np.random.default_rng(11)
    x = (rng.standard_normal((3, 26, 26, 40)) * 3).astype(np.float32)
    b = rng.standard_normal(40).astype(np.float32)
    mean, var = rng.standard_normal(40).astype(np.float32), rng.uniform(0.1, 2, 40).astype(np.float32)
    gamma = rng.uniform(-2, 2, 40).astype(np.float32)
    var_before = var.copy()

    def both(fn):
        c = chain.Chain("oracle"); fn(c, c.create_input(x.shape))
        return _single(x, fn, debug=True), c.run(x)

    for fn in (lambda g, i: g.create_bias_add(i, b),
               lambda g, i: g.create_batch_norm(i, mean, var, gamma, 1e-5),
               lambda g, i: g.create_leaky_relu(i),
               lambda g, i: g.create_max_pool2d(i, [1, 2, 2, 1], [1, 2, 2, 1], 'SAME'),
               lambda g, i: g.create_max_pool2d(i, [1, 2, 2, 1], [1, 1, 1, 1], 'SAME'),
               lambda g, i: g.create_max_pool2d(i, [1, 3, 3, 1], [1, 2, 2, 1], 'VALID')):
        got, ref = both(fn)
        np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(var, var_before)        # unlike the reference, inputs are not written
    k = rng.standard_normal((3, 3, 40, 24)).astype(np.float32)
    got, ref = both(lambda g, i: g.create_conv2d(i, k, [1, 1, 1, 1], 'SAME'))
    assert normalised_error(got, ref) < 1e-5
    got, ref = both(lambda g, i: g.create_conv2d(i, k, [1, 2, 2, 1], 'SAME'))     # strided, the op-level ABI allows it
    assert got.shape == (3, 13, 13, 24) and normalised_error(got, ref) < 1e-5


def test_library_can_be_bound_the_way_the_reference_binds_it(oracle_libs):
    """proj3/dnn_openblas.py:9,180-188,236-242 style: cdll.LoadLibrary, no argtypes/restype, every
    int wrapped in c_int, four extra pad ints on max_pool2d."""
    mylib = ctypes.cdll.LoadLibrary(_lib.LIB_PATH)
    fpt = ctypes.POINTER(ctypes.c_float)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 9, 9, 6)).astype(np.float32)
    k = rng.standard_normal((3, 3, 6, 10)).astype(np.float32)
    xp = np.pad(x, [(0, 0), (1, 1), (1, 1), (0, 0)], 'constant')
    kernel_r = np.ascontiguousarray(k.transpose(2, 0, 1, 3).reshape(-1, 10))
    col = np.zeros((2, 81, 54), np.float32)
    res = np.zeros((2, 9, 9, 10), np.float32)
    mylib.conv2d_mul(xp.ctypes.data_as(fpt), col.ctypes.data_as(fpt), kernel_r.ctypes.data_as(fpt),
                     res.ctypes.data_as(fpt), *map(ctypes.c_int, res.shape), *map(ctypes.c_int, xp.shape[1:]),
                     *map(ctypes.c_int, k.shape[:2]), *map(ctypes.c_int, [1, 1]))
    c = chain.Chain("oracle"); c.create_conv2d(c.create_input(x.shape), k, [1, 1, 1, 1], 'SAME')
    assert normalised_error(res, c.run(x)) < 1e-5
    pp = np.pad(res, [(0, 0), (0, 1), (0, 1), (0, 0)], 'constant', constant_values=np.finfo('float32').min)
    pooled = np.zeros((2, 9, 9, 10), np.float32)
    mylib.max_pool2d(pp.ctypes.data_as(fpt), pooled.ctypes.data_as(fpt), *pooled.shape, *pp.shape[1:], 2, 2, 1, 1, 0, 1, 0, 1)
    c = chain.Chain("oracle"); c.create_max_pool2d(c.create_input(res.shape), [1, 2, 2, 1], [1, 1, 1, 1], 'SAME')
    np.testing.assert_array_equal(pooled, c.run(res))


# ------------------------------------------------------------------------------- fused blocks
def _layer_spec(i, w, pool=True):
    spec = {"kernel": w["kernel"], "padding": 'SAME', "biases": w["biases"]}
    if "gamma" in w:
        spec.update(mean=w["moving_mean"], variance=w["moving_variance"], gamma=w["gamma"], epsilon=1e-5, leaky=True)
    p = POOLS[i] if pool else None
    spec["pool"] = None if p is None else (p[0], p[1], 'SAME')
    return spec


def _oracle_block(i, w, x, pool=True):
    c = chain.Chain("oracle")
    cur = c.create_input(x.shape)
    cur = c.create_conv2d(cur, w["kernel"], [1, 1, 1, 1], 'SAME')
    cur = c.create_bias_add(cur, w["biases"])
    if "gamma" in w:
        cur = c.create_batch_norm(cur, w["moving_mean"], w["moving_variance"], w["gamma"], 1e-5)
        cur = c.create_leaky_relu(cur)
    if pool and POOLS[i] is not None:
        k, s = POOLS[i]
        cur = c.create_max_pool2d(cur, [1, k, k, 1], [1, s, s, 1], 'SAME')
    return c.run(x)


# (layer, batch): batches chosen so the last tile of every tiling is ragged
LAYER_CASES = [(0, 2), (1, 2), (2, 3), (3, 9), (4, 33), (5, 3), (6, 3), (7, 2), (8, 3)]
HW = [416, 208, 104, 52, 26, 13, 13, 13, 13]


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("layer,batch", LAYER_CASES)
def test_each_yolo_block_fused_vs_oracle(oracle_libs, layer, batch, precision):
    w = make_weights(20 + layer, negative_gamma=True)[layer]
    rng = np.random.default_rng(40 + layer)
    x = rng.standard_normal((batch, HW[layer], HW[layer], CHANNELS[layer])).astype(np.float32)
    net = _lib.Net(x.shape, [_layer_spec(layer, w)], precision=precision)
    info = net.block_info(0)
    got = net.forward(x)
    ref = _oracle_block(layer, w, x)
    assert got.shape == ref.shape
    err = normalised_error(got, ref)
    print("layer %d %s: kernel=%s err=%.3g" % (layer, precision, info["kernel"], err))
    if layer == 0:
        assert info["kernel"] == "umma_conv0" and info["pool_fused"]   # 3 channels: patch-resident kernel
    elif layer <= 2:
        assert info["kernel"] == "umma_rowpair" and info["pool_fused"]
    elif layer <= 4:
        assert info["kernel"] == "umma_tiled_pool" and info["pool_fused"]
    else:
        assert info["kernel"] == "umma_im2col"
        assert not info["pool_separate"] and info["pool_fused"] == (layer == 5)     # block 5: the 2x2/1 pool is fused too
    assert err < (2e-3 if precision == "tf32" else 1.2e-2)
    net.close()


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("layer,batch", [(0, 1), (1, 1), (2, 2), (3, 8), (4, 5)])
def test_fused_pool_picks_the_reference_maximum(oracle_libs, layer, batch, precision):
    """Operator-local max-pool identity.  The fused epilogue is asked to also dump the very
    activations it pools (they normally never leave the SM); the oracle's max_pool2d applied to
    that dump must reproduce the fused output bit for bit, hence the same window index."""
    w = make_weights(30 + layer)[layer]
    x = np.random.default_rng(50 + layer).standard_normal((batch, HW[layer], HW[layer], CHANNELS[layer])).astype(np.float32)
    nxt = make_weights(31)[layer + 1]                       # a second block so block 0 stores its real dtype
    net = _lib.Net(x.shape, [_layer_spec(layer, w), _layer_spec(layer + 1, nxt)], precision=precision)
    net.forward(x)
    assert net.block_info(0)["pool_fused"]
    pooled = net.block_output(0)
    prepool = net.debug_prepool(0, (HW[layer], HW[layer]))
    np.testing.assert_array_equal(net.block_output(0), pooled)        # the debug re-run changes nothing
    ref, ref_idx = chain.pool_argmax(prepool, 2, 2)
    np.testing.assert_array_equal(pooled, ref)
    b, h, w_, c = prepool.shape
    win = prepool.reshape(b, h // 2, 2, w_ // 2, 2, c).transpose(0, 1, 3, 5, 2, 4).reshape(b, h // 2, w_ // 2, c, 4)
    got_idx = (win == pooled[..., None]).argmax(-1)       # first position holding the returned value
    np.testing.assert_array_equal(got_idx, ref_idx)
    # and the dumped activations are the block's real ones: they match the oracle's pre-pool tensor
    oracle_pre = _oracle_block(layer, w, x, pool=False)
    assert normalised_error(prepool, oracle_pre) < (2e-3 if precision == "tf32" else 1.2e-2)
    net.close()


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("batch,force_pairs", [(1, False), (3, False), (13, True), (54, True), (54, "no tails")])
def test_fused_stride1_pool_picks_the_reference_maximum(oracle_libs, monkeypatch, batch, force_pairs, precision):
    """Block 5's 2x2 stride-1 SAME pool (max_pool2d_5: pad 0/1 with finfo.min, proj3/dnn_openblas.py:232-235) fused
    into the im2col epilogue over tiles that overlap by W + 1 rows.  Operator-local identity as above: the oracle's
    max_pool2d on the dumped pre-pool activations reproduces the fused output bit for bit, same window index
    (first maximum in (di, dj) raster order, proj3/dnn_openblas.c:7,209-222); and the separate-kernel plan
    (DNNB_NO_POOL1_FUSION) gives the same bits.  With forced CTA pairs batch 13 runs the pair kernel + the
    single-CTA tail, batch 54 (9126 rows = 37 + 3 pair tiles of 228 rows + 6 rows) the pair kernel, the 256 x 128
    pair tail and the single-CTA tail together."""
    layer = 5
    w = make_weights(30 + layer, negative_gamma=(batch == 3))[layer]
    x = np.random.default_rng(50 + layer + batch).standard_normal((batch, 13, 13, CHANNELS[layer])).astype(np.float32)
    nxt = make_weights(31)[layer + 1]
    if force_pairs:
        monkeypatch.setenv("DNNB_PAIR_MIN_TILES", "1")
        if force_pairs is True:          # block 5's short K loop normally gets no tail launches: force them, the variants exist
            monkeypatch.setenv("DNNB_PAIR_FORCE_TAIL", "1")
    net = _lib.Net(x.shape, [_layer_spec(layer, w), _layer_spec(layer + 1, nxt)], precision=precision)
    net.forward(x)
    info = net.block_info(0)
    assert info["pool_fused"] and not info["pool_separate"]
    assert info["kernel"] == ("umma_pair" if force_pairs else "umma_im2col")
    assert info["launches"] == ({13: 2, 54: 3}[batch] if force_pairs is True else 1)
    pooled = net.block_output(0)
    prepool = net.debug_prepool(0, (13, 13))
    np.testing.assert_array_equal(net.block_output(0), pooled)
    ref, ref_idx = chain.pool_argmax(prepool, 2, 1)
    np.testing.assert_array_equal(pooled, ref)
    padded = np.pad(prepool, [(0, 0), (0, 1), (0, 1), (0, 0)], 'constant', constant_values=np.finfo(np.float32).min)
    win = np.stack([padded[:, di:di + 13, dj:dj + 13, :] for di in (0, 1) for dj in (0, 1)], axis=-1)
    got_idx = (win == pooled[..., None]).argmax(-1)
    np.testing.assert_array_equal(got_idx, ref_idx)
    oracle_pre = _oracle_block(layer, w, x, pool=False)
    assert normalised_error(prepool, oracle_pre) < (2e-3 if precision == "tf32" else 1.2e-2)
    second = net.block_output(1)
    net.close()
    monkeypatch.setenv("DNNB_NO_POOL1_FUSION", "1")
    sep = _lib.Net(x.shape, [_layer_spec(layer, w), _layer_spec(layer + 1, nxt)], precision=precision)
    sep.forward(x)
    assert sep.block_info(0)["pool_separate"]
    np.testing.assert_array_equal(sep.block_output(0), pooled)
    np.testing.assert_array_equal(sep.block_output(1), second)
    sep.close()


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("layer,batch", [(3, 1), (4, 1), (5, 1), (4, 3), (5, 2)])
def test_latency_plan_pools_inside_the_conv_kernels(oracle_libs, layer, batch, precision):
    """Small batches under the latency plan give up the spatial tiles (they would be mostly padding) for runs of
    consecutive pixels; the pools are fused all the same -- 2x2/2 over tiles that are whole pairs of image rows
    (blocks 3, 4), 2x2/1 over tiles overlapping by W + 1 rows (block 5) -- including behind the cluster split-K
    reduction, where every CTA ends up with all rows of a column strip.  Operator-local identity as above."""
    w = make_weights(60 + layer, negative_gamma=(batch > 1))[layer]
    x = np.random.default_rng(70 + layer + batch).standard_normal((batch, HW[layer], HW[layer], CHANNELS[layer])).astype(np.float32)
    nxt = make_weights(61)[layer + 1]
    net = _lib.Net(x.shape, [_layer_spec(layer, w), _layer_spec(layer + 1, nxt)], precision=precision, latency=True)
    net.forward(x)
    info = net.block_info(0)
    print("layer %d batch %d %s: %s" % (layer, batch, precision, info))
    assert info["kernel"] == "umma_im2col"
    if batch == 1:
        assert info["pool_fused"] and not info["pool_separate"] and info["launches"] == 1
        assert info["splits"] == 1 or info["cluster_splitk"]
        if layer >= 4:
            assert info["splits"] > 1                  # the weight stream is spread over more CTAs than there are tiles
    else:       # the planner may prefer the workspace split-K (more CTAs), whose pool is a kernel of its own
        assert info["pool_fused"] != info["pool_separate"] and (info["pool_fused"] or not info["cluster_splitk"])
    pooled = net.block_output(0)
    k, st = POOLS[layer]
    if info["pool_fused"]:
        prepool = net.debug_prepool(0, (HW[layer], HW[layer]))
        np.testing.assert_array_equal(net.block_output(0), pooled)
        ref, _ = chain.pool_argmax(prepool, k, st)
        np.testing.assert_array_equal(pooled, ref)
        assert normalised_error(prepool, _oracle_block(layer, w, x, pool=False)) < (2e-3 if precision == "tf32" else 1.2e-2)
    assert normalised_error(pooled, _oracle_block(layer, w, x)) < (2e-3 if precision == "tf32" else 1.2e-2)
    again = net.forward(x).copy()
    np.testing.assert_array_equal(net.forward(x), again)          # reproducible run to run
    net.close()


# ------------------------------------------------------------------------------- end to end
def _yolo_engine(batch, precision, wseed=0, debug=False):
    g = dnn_b200.DnnGraphBuilder()
    build_yolov2tiny(g, [batch, 416, 416, 3], make_weights(wseed))
    return dnn_b200.DnnInferenceEngine(g, debug, precision=precision)


def _class_argmax(out):
    return out.reshape(out.shape[0], 13, 13, 5, 25)[..., 5:].argmax(-1)


@pytest.mark.parametrize("precision", ["tf32", "bf16", "fp32"])
@pytest.mark.parametrize("wseed,iseed", [(0, 100), (1, 101)])
def test_yolo_batch1_against_reference_golden(precision, wseed, iseed):
    gold = np.load(os.path.join(GOLDEN, "yolo_w%d_i%d.npz" % (wseed, iseed)))
    eng = _yolo_engine(1, precision, wseed)
    out = eng.run(make_images(1, seed=iseed))
    ref = gold["out_openblas"]
    assert out.shape == (1, 13, 13, 125) and out.dtype == np.float32
    err = normalised_error(out, ref)
    plan = [p["kernel"] for p in eng.kernel_plan()]
    print("yolo b1 %s w%d: err=%.3g plan=%s" % (precision, wseed, err, plan))
    assert err < TOL[precision]
    if precision != "fp32":
        assert plan == ["umma_conv0"] + ["umma_rowpair"] * 2 + ["umma_tiled_pool"] * 2 + ["umma_im2col"] * 4
    # argmax of detections: identical except on genuine near-ties
    a, r = _class_argmax(out), _class_argmax(ref)
    cls_ref = ref.reshape(1, 13, 13, 5, 25)[..., 5:]
    bad = np.argwhere(a != r)
    bound = 2 * err * np.abs(ref).max()
    for idx in bad:
        logits = cls_ref[tuple(idx)]
        top2 = np.sort(logits)[-2:]
        assert top2[1] - top2[0] <= bound, "argmax flip that is not a near-tie"
    print("   class-argmax mismatches: %d of %d (all near-ties)" % (len(bad), a.size))
    if precision == "fp32":
        assert len(bad) == 0
    eng.close()


def test_yolo_block_outputs_layerwise_against_oracle(oracle_libs):
    """Layer-wise comparison at block granularity (the reference's layer_N convention, restricted
    to the tensors that still exist after fusion)."""
    weights = make_weights(0)
    x = make_images(1, seed=100)
    _, nodes = chain.yolo_forward(weights, x, "oracle", keep=True)
    eng = _yolo_engine(1, "tf32")
    eng.run(x)
    ends = [5, 10, 15, 20, 25, 30, 34, 38, 40]           # index of each block's last node
    for blk, (node, res) in enumerate(eng.block_results().items()):
        ref = nodes[ends[blk]]
        assert res.shape == ref.shape
        assert normalised_error(res, ref) < 5e-3, "block %d" % blk
    # the same through layer_compare.compare_blocks: goldens on disk in the proj1 numbering (layer_0 = conv2d_0)
    import layer_compare, tempfile
    with tempfile.TemporaryDirectory() as d:
        for i, t in enumerate(nodes[1:]):
            np.save(os.path.join(d, "layer_%d.npy" % i), t)
        rows = layer_compare.compare_blocks(eng, d, quiet=True)
    assert [r["index"] for r in rows] == [e - 1 for e in ends] and rows[0]["name"] == "max_pool2d_0"
    assert max(r["normalised"] for r in rows) < 5e-3
    eng.close()


def test_batch_results_do_not_depend_on_batch_size(oracle_libs):
    """Image i of a batch-5 run is bit-identical to a batch-1 run of image i (per-image math does
    not depend on how the batch is tiled): the property multi-GPU sharding relies on."""
    x = make_images(5, seed=200)
    eng5 = _yolo_engine(5, "tf32")
    out5 = eng5.run(x).copy()
    eng1 = _yolo_engine(1, "tf32")
    for i in (0, 3, 4):
        np.testing.assert_array_equal(eng1.run(x[i:i + 1])[0], out5[i])
    ref = chain.yolo_forward(make_weights(0), x[4:5], "oracle")
    assert normalised_error(out5[4:5], ref) < TOL["tf32"]
    out5b = eng5.run(x)
    np.testing.assert_array_equal(out5, out5b)            # deterministic
    eng5.close(); eng1.close()


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_cta_pair_kernel_equals_the_single_cta_kernel_bit_for_bit(oracle_libs, monkeypatch, precision):
    """Blocks 5-7 on CTA pairs (tcgen05.mma.cta_group::2, conv_pair.cuh).  The planner only uses pairs
    at large batches; DNNB_PAIR_MIN_TILES=1 forces them at batch 13 (2197 rows: 8 whole 256-row pair tiles
    plus a ragged tail launch), both tile heights.  Every K block is summed in the same order as in
    the single-CTA kernel, so the outputs must be identical, and within tolerance of the oracle."""
    B = 13
    x = make_images(B, seed=400)

    def run(env):
        for k in ("DNNB_NO_PAIR", "DNNB_PAIR_MT", "DNNB_PAIR_MIN_TILES"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = _yolo_engine(B, precision)
        out = eng.run(x).copy()
        plan = [p["kernel"] for p in eng.kernel_plan()]
        blocks = [eng._net.block_output(i) for i in (5, 6, 7)]
        eng.close()
        return out, plan, blocks

    ref_out, ref_plan, ref_blocks = run({"DNNB_NO_PAIR": "1"})
    assert "umma_pair" not in ref_plan
    for mt in ("1", "2"):
        out, plan, blocks = run({"DNNB_PAIR_MIN_TILES": "1", "DNNB_PAIR_MT": mt})
        assert plan[5:8] == ["umma_pair"] * 3, plan
        for a, b in zip(blocks, ref_blocks):
            np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(out, ref_out)
    ref = chain.yolo_forward(make_weights(0), x[12:13], "oracle")
    assert normalised_error(ref_out[12:13], ref) < TOL[precision]


def test_full_size_batch256_properties(oracle_libs):
    """BASELINE configs[2] size.  Size-independent checks: sampled images match the batch-1
    engine bit for bit and the oracle within tolerance; no NaN/Inf anywhere."""
    B = 256
    x = make_images(B, seed=1000)
    eng = _yolo_engine(B, "tf32")
    out = eng.run(x)
    assert out.shape == (B, 13, 13, 125) and np.isfinite(out).all()
    assert [p["kernel"] for p in eng.kernel_plan()][5:8] == ["umma_pair"] * 3     # CTA pairs at this size
    eng1 = _yolo_engine(1, "tf32")
    for i in (0, 127, 255):
        np.testing.assert_array_equal(eng1.run(x[i:i + 1])[0], out[i])
    ref = chain.yolo_forward(make_weights(0), x[255:256], "oracle")
    assert normalised_error(out[255:256], ref) < TOL["tf32"]
    eng.close(); eng1.close()


@pytest.mark.parametrize("hw,batch", [(320, 2), (352, 1), (96, 3)])
def test_other_input_sizes(oracle_libs, hw, batch):
    """Nothing in the kernels is tied to 416: the YOLOv2-tiny stack on other multiples of 32."""
    weights = make_weights(2)
    x = make_images(batch, seed=300, hw=hw)
    g = dnn_b200.DnnGraphBuilder()
    build_yolov2tiny(g, [batch, hw, hw, 3], weights)
    eng = dnn_b200.DnnInferenceEngine(g, False, precision="tf32")
    out = eng.run(x)
    plan = [p["kernel"] for p in eng.kernel_plan()]
    c = chain.Chain("oracle"); build_yolov2tiny(c, x.shape, weights)
    ref = c.run(x)
    assert out.shape == ref.shape == (batch, hw // 32, hw // 32, 125)
    err = normalised_error(out, ref)
    print("hw %d: err=%.3g plan=%s" % (hw, err, plan))
    assert err < TOL["tf32"]
    assert all(k.startswith("umma") for k in plan)
    eng.close()


def test_batch_sharding_over_two_gpus_equals_one_gpu():
    """the G-GPU run is bit-identical to the 1-GPU run of the same build on the same inputs."""
    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import sharded
    weights = make_weights(0)
    spec = []
    for i, w in enumerate(weights):
        blk = {"kernel": w["kernel"], "padding": "SAME", "biases": w["biases"]}
        if "gamma" in w:
            blk.update(mean=w["moving_mean"], variance=w["moving_variance"], gamma=w["gamma"], epsilon=1e-5, leaky=True)
        blk["pool"] = None if POOLS[i] is None else (POOLS[i][0], POOLS[i][1], "SAME")
        spec.append(blk)
    total = 7                                           # ragged: shards of 4 and 3 images
    x = make_images(total, seed=400)
    one = _lib.Net(x.shape, spec, precision="tf32", device=0)
    ref = one.forward(x)
    eng = sharded.MultiGpuEngine(lambda dev, b: _lib.Net((b, 416, 416, 3), spec, precision="tf32", device=dev), total, [0, 1])
    hin, hout = _lib.PinnedArray(x.shape), _lib.PinnedArray(ref.shape)
    hin.array[...] = x
    out = eng.run(hin.array, hout.array)
    np.testing.assert_array_equal(out, ref)
    eng.close(); one.close(); hin.free(); hout.free()


def test_sharded_engine_through_the_drop_in_api_on_one_device():
    """DnnInferenceEngine(devices=[...]) is the batch-sharded driver behind the reference's API.  With the same
    device listed twice the whole sharded path (ragged slices, one resident network per slice, async forwards,
    result assembled in one host array) runs on a 1-GPU box; slices must land where a single network puts them."""
    weights = make_weights(0)
    x = make_images(5, seed=410)
    g = dnn_b200.DnnGraphBuilder()
    build_yolov2tiny(g, list(x.shape), weights)
    one = dnn_b200.DnnInferenceEngine(g, False, precision="tf32", device=0)
    ref = one.run(x).copy()
    one.close()
    g2 = dnn_b200.DnnGraphBuilder()
    build_yolov2tiny(g2, list(x.shape), weights)
    eng = dnn_b200.DnnInferenceEngine(g2, False, precision="tf32", devices=[0, 0])
    out = eng.run(x)
    assert [n.in_shape[0] for n in eng._net.nets] == [3, 2]
    np.testing.assert_array_equal(out, ref)
    assert g2.out_node.result is out
    assert all(p["kernel"].startswith("umma") for p in eng.kernel_plan())
    eng.close()


def test_async_entry_points_reject_buffers_they_would_overrun():
    spec = [{"kernel": np.zeros((3, 3, 3, 16), np.float32), "padding": "SAME", "biases": np.zeros(16, np.float32)}]
    net = _lib.Net((2, 16, 16, 3), spec, precision="tf32", device=0)
    out = np.zeros(net.out_shape, np.float32)
    with pytest.raises(TypeError):
        net.forward_async(np.zeros((2, 16, 16, 3), np.float64), out)
    with pytest.raises(ValueError):
        net.forward_async(np.zeros((4, 16, 16, 3), np.float32)[::2], out)
    with pytest.raises(ValueError):
        net.forward_async(np.zeros((2, 16, 16, 3), np.float32), np.zeros((1,) + net.out_shape[1:], np.float32))
    with pytest.raises(TypeError):
        net.forward_u8_async(np.zeros((2, 16, 16, 3), np.float32), out)
    net.close()


def test_unknown_environment_switch_fails_finalize(monkeypatch):
    spec = [{"kernel": np.zeros((3, 3, 3, 16), np.float32), "padding": "SAME", "biases": np.zeros(16, np.float32)}]
    monkeypatch.setenv("DNNB_NO_PIAR", "1")                       # a typo must not silently run the default plan
    with pytest.raises(RuntimeError, match="unknown environment switch DNNB_NO_PIAR"):
        _lib.Net((1, 16, 16, 3), spec)
    monkeypatch.delenv("DNNB_NO_PIAR")
    monkeypatch.setenv("DNNB_PAIR_MT", "three")
    with pytest.raises(RuntimeError, match="bad value for DNNB_PAIR_MT"):
        _lib.Net((1, 16, 16, 3), spec)
    monkeypatch.delenv("DNNB_PAIR_MT")
    monkeypatch.setenv("DNNB_NO_PAIR", "1")
    monkeypatch.setenv("DNNB_PAIR_MIN_TILES", "1")
    net = _lib.Net((1, 16, 16, 3), spec, latency=True)
    assert net.knobs() == {"DNNB_NO_PAIR": "1", "DNNB_PAIR_MIN_TILES": "1", "set_latency": "1"}
    net.close()
    monkeypatch.delenv("DNNB_NO_PAIR"); monkeypatch.delenv("DNNB_PAIR_MIN_TILES")
    net = _lib.Net((1, 16, 16, 3), spec)
    assert net.knobs() == {}
    net.close()


def test_latency_widths_pin_the_plan_and_tuning_measures_it():
    """dnnb_net_set_latency_widths / latency='tune': a width per block decides tiles x K slices; whatever the choice the
    result stays within the reordering noise of the default plan, and the measured plan is never slower than it."""
    x = make_images(1, seed=100)
    spec = [_layer_spec(i, w) for i, w in enumerate(make_weights(0))]
    ref_net = _lib.Net(x.shape, spec, precision="tf32", latency=True)
    ref = ref_net.forward(x)
    base_bn = [ref_net.block_info(i)["bn_tile"] for i in range(9)]
    ref_net.close()
    widths = [0, 0, 0, 128, 128, 128, 256, 256, 128]
    net = _lib.Net(x.shape, spec, precision="tf32", latency=True, latency_widths=widths)
    got = net.forward(x)
    assert [net.block_info(i)["bn_tile"] for i in range(3, 9)] == [128, 128, 128, 256, 256, 128] != base_bn[3:]
    assert net.knobs().get("set_latency_widths") == "0/0/0/128/128/128/256/256/128"
    assert normalised_error(got, ref) < 2e-3
    net.close()
    with pytest.raises(RuntimeError):
        _lib.Net(x.shape, spec, precision="tf32", latency=True, latency_widths=[0, 0, 0, 48])
    tuned = _lib.tune_latency_widths(x.shape, spec, "tf32", iters=20, sweeps=1)
    assert len(tuned["widths"]) == 9 and tuned["tuned_us"] <= tuned["modelled_us"] and tuned["modelled_us"] < 200
    g = dnn_b200.DnnGraphBuilder()
    build_yolov2tiny(g, [1, 416, 416, 3], make_weights(0))
    eng = dnn_b200.DnnInferenceEngine(g, False, precision="tf32", latency="tune")
    out = eng.run(x)
    assert eng.latency_widths is not None and normalised_error(out, ref) < 2e-3
    eng.close()


def test_phase_timeline_of_the_latency_plan():
    """dnnb_net_debug_timeline: per-CTA %globaltimer stamps of the generic kernel (tools/batch1_timeline.py).  Every CTA
    of an armed block stamps its phases in order, and arming does not change the result."""
    x = make_images(1, seed=100)
    eng = _yolo_engine(1, "tf32")
    eng.latency = True
    ref = eng.run(x).copy()
    net = eng._net
    info = net.block_info(7)
    assert info["kernel"] == "umma_im2col" and info["splits"] > 1
    assert net.debug_timeline(7) == info["grid"]
    out = eng.run(x)
    np.testing.assert_array_equal(out, ref)
    t = net.debug_timeline(7, read=True).astype(np.int64)
    assert t.shape == (info["grid"], _lib.TIMELINE_SLOTS) and (t[:, :6] > 0).all()    # slots 6-7 belong to the CTAs that reduce and store
    assert (t[:, 7] > 0).any()
    for row in t[:, [0, 1, 2, 3, 4, 5, 8, 6, 9, 7]]:              # in phase order; 8-9 are the split-K sub-phases
        seen = row[row > 0]
        assert (np.diff(seen) >= 0).all()                           # entry <= prologue <= ... <= stored, per CTA
    assert (t.max() - t[:, 0].min()) < 200_000                      # a block is tens of microseconds, not milliseconds
    with pytest.raises(RuntimeError):
        net.debug_timeline(0)                                      # block 0 does not run the generic kernel
    eng.close()


def test_valid_padding_and_odd_pool_take_the_cuda_core_path(oracle_libs):
    """A block the tcgen05 kernels do not cover (VALID conv, odd sizes, 3x3/2 pool) still runs on the GPU."""
    rng = np.random.default_rng(8)
    x = rng.standard_normal((2, 15, 17, 6)).astype(np.float32)
    k = rng.standard_normal((3, 3, 6, 10)).astype(np.float32)
    b = rng.standard_normal(10).astype(np.float32)
    net = _lib.Net(x.shape, [{"kernel": k, "padding": 'VALID', "biases": b, "leaky": True, "pool": (3, 2, 'VALID')}], precision="tf32")
    assert net.block_info(0)["kernel"] == "simt"
    c = chain.Chain("oracle")
    cur = c.create_conv2d(c.create_input(x.shape), k, [1, 1, 1, 1], 'VALID')
    cur = c.create_leaky_relu(c.create_bias_add(cur, b))
    c.create_max_pool2d(cur, [1, 3, 3, 1], [1, 2, 2, 1], 'VALID')
    ref = c.run(x)
    got = net.forward(x)
    assert got.shape == ref.shape and normalised_error(got, ref) < 1e-5
    net.close()


def test_latency_mode_split_k_stays_within_reordering_noise(oracle_libs, monkeypatch):
    """DNNB_LATENCY_MODE=1: split-K over co-resident CTAs for small batches.  Reproducible run to run
    (slices are summed in order); not bit-identical to the default plan, whose K sum is cut differently."""
    x = make_images(1, seed=100)
    base = _yolo_engine(1, "tf32")
    ref = base.run(x).copy()
    base.close()
    monkeypatch.setenv("DNNB_LATENCY_MODE", "1")
    eng = _yolo_engine(1, "tf32")
    out = eng.run(x).copy()
    out2 = eng.run(x).copy()
    eng.close()
    gold = np.load(os.path.join(GOLDEN, "yolo_w0_i100.npz"))["out_openblas"]
    # a re-ordered fp32 sum can flip a TF32 rounding of an activation (2^-11 relative), which then propagates:
    # the two engines differ by about as much as either differs from the reference
    np.testing.assert_array_equal(out2, out)
    assert normalised_error(out, ref) < 2e-3
    assert normalised_error(out, gold) < TOL["tf32"]


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_latency_mode_graph_replay_equals_direct_launches(monkeypatch, precision):
    """DNNB_LATENCY_MODE=1 also chains the kernels by programmatic dependent launch and replays the
    device-resident forward (dnnb_net_forward_device) from a CUDA graph from the second call on the same
    pointers.  Every replay must reproduce the directly launched result bit for bit -- the split-K
    counters reset themselves on the device -- also after the input buffer's contents change, and the
    graph-free, PDL-free launch sequence (DNNB_NO_GRAPH, DNNB_NO_PDL) must give the same bits."""
    import torch
    monkeypatch.setenv("DNNB_LATENCY_MODE", "1")
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    xs = [torch.from_numpy(make_images(1, seed=s)).to(dev) for s in (100, 101)]
    x = torch.empty_like(xs[0])
    y = torch.empty((1, 13, 13, 125), device=dev)

    def run_all(calls):
        eng = _yolo_engine(1, precision)
        eng.kernel_plan()                       # builds the resident network
        outs = []
        with torch.cuda.stream(stream):
            for i in range(calls):
                x.copy_(xs[i % 2], non_blocking=True)
                eng._net.forward_device(x.data_ptr(), y.data_ptr(), stream.cuda_stream)
                stream.synchronize()
                outs.append(y.cpu().numpy().copy())
        eng.close()
        return outs

    outs = run_all(12)                          # call 0 direct; call 1 (the buffer pair again) is captured and replayed, as are 2..
    monkeypatch.setenv("DNNB_NO_GRAPH", "1")
    monkeypatch.setenv("DNNB_NO_PDL", "1")
    plain = run_all(2)
    for i, o in enumerate(outs):
        np.testing.assert_array_equal(o, plain[i % 2], err_msg="call %d" % i)
    assert not np.array_equal(plain[0], plain[1])
    gold = np.load(os.path.join(GOLDEN, "yolo_w0_i100.npz"))["out_openblas"]
    assert normalised_error(outs[0], gold) < TOL[precision]


def test_latency_keyword_equals_the_environment_switch(monkeypatch):
    """DnnInferenceEngine(..., latency=True) plans like DNNB_LATENCY_MODE=1 without touching the caller's
    environment; latency=False overrides the variable the other way."""
    x = make_images(1, seed=100)

    def run(**kw):
        g = dnn_b200.DnnGraphBuilder()
        build_yolov2tiny(g, [1, 416, 416, 3], make_weights(0))
        eng = dnn_b200.DnnInferenceEngine(g, False, precision="tf32", **kw)
        out = eng.run(x).copy()
        tiles = [p["tiles"] for p in eng.kernel_plan()]
        eng.close()
        return out, tiles

    monkeypatch.delenv("DNNB_LATENCY_MODE", raising=False)
    out_default, tiles_default = run()
    out_kw, tiles_kw = run(latency=True)
    assert "DNNB_LATENCY_MODE" not in os.environ
    monkeypatch.setenv("DNNB_LATENCY_MODE", "1")
    out_env, tiles_env = run()
    out_off, tiles_off = run(latency=False)
    assert os.environ["DNNB_LATENCY_MODE"] == "1"
    assert tiles_kw == tiles_env and tiles_kw != tiles_default and tiles_off == tiles_default
    assert tiles_kw[7] > 100 > tiles_default[7]          # block 7 is split over (nearly) all SMs
    np.testing.assert_array_equal(out_kw, out_env)
    np.testing.assert_array_equal(out_off, out_default)


def test_uint8_frames_match_numpy_preprocessing():
    """proj3/__init__.py:8-12 minus the resize.  The device-side /255 + BGR->RGB must be
    bit-identical to numpy's, so the network output equals the float32 path's exactly."""
    rng = np.random.default_rng(77)
    frames = rng.integers(0, 256, size=(3, 416, 416, 3), dtype=np.uint8)
    x = np.asarray((frames / 255.)[:, :, :, ::-1], dtype=np.float32)      # resize_input's arithmetic
    g = dnn_b200.DnnGraphBuilder()
    build_yolov2tiny(g, [3, 416, 416, 3], make_weights(0))
    eng = dnn_b200.DnnInferenceEngine(g, False, precision="tf32")
    ref = eng.run(x).copy()
    got = eng._net.forward_u8(frames, bgr=True)
    np.testing.assert_array_equal(got, ref)
    got_rgb = eng._net.forward_u8(np.ascontiguousarray(frames[..., ::-1]), bgr=False)
    np.testing.assert_array_equal(got_rgb, ref)
    got2 = eng._net.forward_u8(frames, bgr=True)      # second call: the other half of the double buffers
    np.testing.assert_array_equal(got2, ref)
    eng.close()


def test_debug_mode_runs_node_by_node_and_dumps_layers(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    gold = np.load(os.path.join(GOLDEN, "mini_net.npz"))
    weights = make_weights(5, channels=[3, 8, 16, 10], ksizes=[3, 3, 1], negative_gamma=True)
    x = make_images(1, seed=55, hw=13)
    g = dnn_b200.DnnGraphBuilder()
    cur = g.create_input(list(x.shape))
    for i, w in enumerate(weights):
        cur = g.create_conv2d(cur, w["kernel"], [1, 1, 1, 1], 'SAME')
        cur = g.create_bias_add(cur, w["biases"])
        if i < 2:
            cur = g.create_batch_norm(cur, w["moving_mean"], w["moving_variance"], w["gamma"], 1e-5)
            cur = g.create_leaky_relu(cur)
            cur = g.create_max_pool2d(cur, [1, 2, 2, 1], [1, 2 - i, 2 - i, 1], 'SAME')
    g.set_out_node(cur)
    out = dnn_b200.DnnInferenceEngine(g, True).run(x)
    n_nodes = len(gold.files)
    for i in range(n_nodes):
        got = np.load(tmp_path / "intermediate" / ("layer_%d.npy" % i))
        np.testing.assert_allclose(got, gold["node_%02d" % i], rtol=1e-5, atol=1e-5, err_msg="layer %d" % i)
    np.testing.assert_allclose(out, gold["node_%02d" % (n_nodes - 1)], rtol=1e-5, atol=1e-5)
    # the reference's own way of checking the dump: npc_cmp_print against a directory of golden layer_N.npy
    import layer_compare
    os.makedirs(tmp_path / "golden")
    for i in range(n_nodes):
        np.save(tmp_path / "golden" / ("layer_%d.npy" % i), gold["node_%02d" % i])
    rows = layer_compare.compare_dirs(str(tmp_path / "intermediate"), str(tmp_path / "golden"), quiet=True)
    assert len(rows) == n_nodes and max(r["normalised"] for r in rows) < 1e-5
    # the same graph, fused (small odd shapes -> CUDA-core kernels + separate pools): same answer
    fused = dnn_b200.DnnInferenceEngine(g, False, precision="fp32").run(x)
    np.testing.assert_allclose(fused, out, rtol=1e-4, atol=1e-4)
    fused_tf32 = dnn_b200.DnnInferenceEngine(g, False, precision="tf32").run(x)
    assert normalised_error(fused_tf32, out) < 5e-3


def test_errors_are_reported_not_swallowed():
    with pytest.raises(RuntimeError, match="device 99"):
        _lib.Net((1, 8, 8, 4), [{"kernel": np.zeros((3, 3, 4, 4), np.float32)}], device=99)
    with pytest.raises(RuntimeError, match="input channels"):
        _lib.Net((1, 8, 8, 4), [{"kernel": np.zeros((3, 3, 5, 4), np.float32)}])


def test_tensor_pipe_peak_probe_runs():
    tf32 = _lib.measure_mma_peak("tf32")
    bf16 = _lib.measure_mma_peak("bf16")
    print("tcgen05 peak: tf32 %.0f TFLOP/s, bf16 %.0f TFLOP/s" % (tf32, bf16))
    assert 300 < tf32 < 1400 and 600 < bf16 < 2800
