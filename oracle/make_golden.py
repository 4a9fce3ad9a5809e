"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running THE REFERENCE ITSELF:
the reference's own host layer (proj3/dnn_openblas.py, proj3/dnn_avx.py imported from
/root/reference) over its own C sources compiled unmodified (oracle/build.py).

Run in the build container (needs /root/reference):   python oracle/make_golden.py
The fixtures are small and committed; the GPU box and the test-suite only read them.

Reference quirks honoured: batch 1, one run() per graph build, a
private copy of moving_variance per build, float32 everywhere, cwd-relative LoadLibrary,
an ./intermediate directory for the OpenBLAS engine's unconditional np.save.
"""
import contextlib
import io
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/cs492-projects/proj3"
sys.path.insert(0, os.path.join(ROOT, "dnn-inference-engine_b200"))
sys.path.insert(0, HERE)

import build as oracle_build  # noqa: E402
from synth import build_yolov2tiny, make_images, make_weights  # noqa: E402

SAMPLE = 64  # values kept per node (strided), besides the statistics


def import_reference(workdir):
    """chdir into a scratch dir holding the two reference .so files, import the reference modules."""
    for so in oracle_build.build_ref():
        shutil.copy(so, workdir)
    os.makedirs(os.path.join(workdir, "intermediate"), exist_ok=True)
    os.chdir(workdir)
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import dnn_avx
        import dnn_openblas
    return dnn_openblas, dnn_avx


def run_reference(mod, weights, image):
    """One fresh graph, one run (App. B1/B4).  Returns the per-node results in execution order."""
    with contextlib.redirect_stdout(io.StringIO()):
        g = mod.DnnGraphBuilder()
        nodes = build_yolov2tiny(g, list(image.shape), weights, copy_variance=True)
        eng = mod.DnnInferenceEngine(g, False)
        out = eng.run(image)
    assert out is nodes[-1].result
    return [np.array(n.result, dtype=np.float32, copy=True) for n in nodes[1:]]


def node_digest(results):
    sums = np.array([r.astype(np.float64).sum() for r in results])
    absmax = np.array([np.abs(r).max() for r in results], dtype=np.float32)
    samples = np.stack([r.ravel()[:: max(1, r.size // SAMPLE)][:SAMPLE] for r in results])
    shapes = np.array([list(r.shape) for r in results], dtype=np.int64)
    return sums, absmax, samples.astype(np.float32), shapes


def small_ops(mod_ob, rng):
    """Op-level known-answer cases through the reference's node classes."""
    out = {}

    def graph(build):
        with contextlib.redirect_stdout(io.StringIO()):
            g = mod_ob.DnnGraphBuilder()
            x_shape, fn = build
            inp = g.create_input(list(x_shape))
            node = fn(g, inp)
            g.set_out_node(node)
            return g, inp, node

    def run(name, x, fn):
        g, inp, node = graph((x.shape, fn))
        with contextlib.redirect_stdout(io.StringIO()):
            res = mod_ob.DnnInferenceEngine(g, False).run(x)
        out[name + "_x"] = x
        out[name + "_y"] = np.array(res, copy=True)

    f32 = np.float32
    x = rng.standard_normal((1, 6, 7, 5)).astype(f32)
    k3 = rng.standard_normal((3, 3, 5, 4)).astype(f32); out["conv3_k"] = k3
    run("conv3_same", x, lambda g, i: g.create_conv2d(i, k3, [1, 1, 1, 1], 'SAME'))
    run("conv3_valid", x, lambda g, i: g.create_conv2d(i, k3, [1, 1, 1, 1], 'VALID'))
    k1 = rng.standard_normal((1, 1, 5, 3)).astype(f32); out["conv1_k"] = k1
    run("conv1_same", x, lambda g, i: g.create_conv2d(i, k1, [1, 1, 1, 1], 'SAME'))
    b = rng.standard_normal(5).astype(f32); out["bias_b"] = b
    run("bias", x, lambda g, i: g.create_bias_add(i, b))
    mean = rng.standard_normal(5).astype(f32); var = rng.uniform(0.5, 1.5, 5).astype(f32)
    gamma = rng.uniform(-1.5, 1.5, 5).astype(f32)
    out["bn_mean"], out["bn_var"], out["bn_gamma"] = mean, var.copy(), gamma
    run("bn", x, lambda g, i: g.create_batch_norm(i, mean, var.copy(), gamma, 1e-5))
    run("leaky", x, lambda g, i: g.create_leaky_relu(i))
    xe = rng.standard_normal((1, 6, 8, 5)).astype(f32)
    run("pool_s2_even", xe, lambda g, i: g.create_max_pool2d(i, [1, 2, 2, 1], [1, 2, 2, 1], 'SAME'))
    run("pool_s2_odd", x, lambda g, i: g.create_max_pool2d(i, [1, 2, 2, 1], [1, 2, 2, 1], 'SAME'))
    xo = rng.standard_normal((1, 5, 5, 3)).astype(f32)
    run("pool_s1_same", xo, lambda g, i: g.create_max_pool2d(i, [1, 2, 2, 1], [1, 1, 1, 1], 'SAME'))
    xt = np.round(rng.standard_normal((1, 6, 6, 4)) * 2).astype(f32) / 2  # many exact ties
    run("pool_ties", xt, lambda g, i: g.create_max_pool2d(i, [1, 2, 2, 1], [1, 2, 2, 1], 'SAME'))
    return out


def postproc_golden(ref_yolo):
    """The reference's own `postprocessing` (proj3/yolov2tiny.py:94) on seeded head outputs."""
    import postproc as pp
    out = {}
    stored = pp.synth_predictions(8, seed=11)                 # kept whole in the fixture
    seeded = pp.synth_predictions(40, seed=12, spread=1.3)    # regenerated from the seed by the tests
    out["pred_stored"] = stored
    out["seeded_seed"], out["seeded_batch"], out["seeded_spread"] = 12, 40, 1.3
    out["seeded_checksum"] = np.float64(seeded.astype(np.float64).sum())
    rows, counts = [], []
    for p in list(stored) + list(seeded):
        boxes = ref_yolo.postprocessing(p)
        counts.append(len(boxes))
        for name, lt, rb, color in boxes:
            rows.append([pp.CLASS_NAMES.index(name), lt[0], lt[1], rb[0], rb[1]])
    out["det_count"] = np.array(counts, dtype=np.int64)
    out["det_rows"] = np.array(rows, dtype=np.int64).reshape(-1, 5)
    colors = np.zeros((pp.CLASSES, 3))
    for c in range(pp.CLASSES):                               # one confident box of each class -> its colour
        p = np.full((13, 13, 5, 25), -20, np.float32)
        p[6, 6, 0, :5] = [0, 0, 0, 0, 10]
        p[6, 6, 0, 5 + c] = 20
        (name, lt, rb, color), = ref_yolo.postprocessing(p.reshape(13, 13, 125))
        assert name == pp.CLASS_NAMES[c]
        colors[c] = color
        if c == 0:
            out["single_box"] = np.array([lt[0], lt[1], rb[0], rb[1]], dtype=np.int64)
    out["colors"] = colors
    return out


def resize_golden():
    """cv2.resize (what proj3/__init__.py:9 calls) on seeded frames; small target sizes keep the
    fixture small, one 416x416 case is stored as a checksum and a strided sample."""
    import cv2
    rng = np.random.default_rng(21)
    out = {"cv2_version": np.array(cv2.__version__)}
    cases = [(50, 70, 96), (200, 150, 96), (96, 96, 96), (97, 33, 96), (48, 48, 96), (3, 5, 32), (1, 1, 32), (131, 207, 64)]
    out["cases"] = np.array(cases, dtype=np.int64)
    for i, (h, w, d) in enumerate(cases):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if i == 1:
            img[:, :, :] = (np.arange(w)[None, :, None] * 255 // max(1, w - 1)).astype(np.uint8)    # a smooth ramp
        out["src_%d" % i] = img
        out["dst_%d" % i] = cv2.resize(img, (d, d))
    big = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    r = cv2.resize(big, (416, 416))
    out["big_src"], out["big_sum"], out["big_sample"] = big, np.int64(r.astype(np.int64).sum()), r[::13, ::7].copy()
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    golden = os.path.join(ROOT, "tests", "golden")
    os.makedirs(golden, exist_ok=True)
    cwd = os.getcwd()
    work = tempfile.mkdtemp(prefix="golden_")
    try:
        mod_ob, mod_avx = import_reference(work)
        np.savez_compressed(os.path.join(golden, "ops_small.npz"), **small_ops(mod_ob, np.random.default_rng(7)))
        np.savez_compressed(os.path.join(golden, "resize.npz"), **resize_golden())
        import yolov2tiny as ref_yolo      # imports dnn_openblas: needs the cwd set up above
        pg = postproc_golden(ref_yolo)
        np.savez_compressed(os.path.join(golden, "postproc.npz"), **pg)
        print("postproc: %d images, %d boxes" % (len(pg["det_count"]), len(pg["det_rows"])))
        for wseed, iseed in [(0, 100), (1, 101)]:
            weights = make_weights(wseed)
            image = make_images(1, seed=iseed)
            res_ob = run_reference(mod_ob, weights, image)
            res_avx = run_reference(mod_avx, weights, image)
            sums, absmax, samples, shapes = node_digest(res_ob)
            np.savez_compressed(
                os.path.join(golden, "yolo_w%d_i%d.npz" % (wseed, iseed)),
                weight_seed=wseed, image_seed=iseed,
                out_openblas=res_ob[-1], out_avx=res_avx[-1],
                node_sum=sums, node_absmax=absmax, node_sample=samples, node_shape=shapes)
            d = np.abs(res_ob[-1] - res_avx[-1]).max() / np.abs(res_ob[-1]).max()
            print("w%d i%d: out absmax %.4f, openblas-vs-avx normalised diff %.3g" %
                  (wseed, iseed, np.abs(res_ob[-1]).max(), d))
        # a small 3-block network (2 pools incl. the stride-1 SAME pool) with EVERY node kept whole
        chs, ks = [3, 8, 16, 10], [3, 3, 1]
        weights = make_weights(5, channels=chs, ksizes=ks, negative_gamma=True)
        image = make_images(1, seed=55, hw=13)
        with contextlib.redirect_stdout(io.StringIO()):
            g = mod_ob.DnnGraphBuilder()
            cur = g.create_input(list(image.shape)); nodes = []
            for i, w in enumerate(weights):
                cur = g.create_conv2d(cur, w["kernel"], [1, 1, 1, 1], 'SAME'); nodes.append(cur)
                cur = g.create_bias_add(cur, w["biases"]); nodes.append(cur)
                if i < 2:
                    cur = g.create_batch_norm(cur, w["moving_mean"], w["moving_variance"].copy(), w["gamma"], 1e-5); nodes.append(cur)
                    cur = g.create_leaky_relu(cur); nodes.append(cur)
                    cur = g.create_max_pool2d(cur, [1, 2, 2, 1], [1, 2 - i, 2 - i, 1], 'SAME'); nodes.append(cur)
            g.set_out_node(cur)
            mod_ob.DnnInferenceEngine(g, False).run(image)
        np.savez_compressed(os.path.join(golden, "mini_net.npz"),
                            **{"node_%02d" % i: np.array(n.result, copy=True) for i, n in enumerate(nodes)})
        print("wrote", sorted(os.listdir(golden)))
    finally:
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
