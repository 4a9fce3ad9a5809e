"""Differential test on the GPU box (pytest -m gpu): seeded random stacks of the five reference ops -- odd and even
image sizes, channel counts that are not multiples of anything, 1x1 and 3x3 kernels, SAME and VALID padding, both
2x2 pools of the YOLO graph or none, batch-norm with negative gamma or no batch-norm at all, TF32 and BF16, default
and latency plan -- through the fused C-ABI network and, node for node, through the CPU oracle (oracle/).  Whatever
kernel the planner picks for a shape (tcgen05 tiled / im2col / row-pair / conv0 / CUDA-core fallback, split-K in
either form), the result has to be the reference's within the precision's tolerance (tests/test_gpu_parity.py: TOL).
"""
import numpy as np
import pytest

import _lib
import chain

pytestmark = pytest.mark.gpu

TOL = {"tf32": 1e-2, "bf16": 3e-2}


def _random_case(seed):
    rng = np.random.default_rng(7000 + seed)
    batch = int(rng.choice([1, 1, 2, 3, 5, 9]))
    h, w = int(rng.integers(5, 45)), int(rng.integers(5, 45))
    friendly = seed % 2 == 0                            # channel counts the tcgen05 kernels take (whole 128-byte K chunks)
    if friendly:
        h, w = [(13, 13), (26, 26), (2 * (h // 2 + 1), 2 * (w // 2 + 1)), (h, w)][int(rng.integers(0, 4))]
    ic = int(rng.choice([32, 64, 128]) if friendly else rng.choice([1, 3, 4, 6, 8, 12, 16, 20, 32, 40]))
    blocks, shape = [], (batch, h, w, ic)
    for _ in range(int(rng.integers(1, 4))):
        ksz = int(rng.choice([1, 3, 3]))
        oc = int(rng.choice([64, 64, 128, 128, 125, 256, 192]) if friendly else rng.choice([1, 5, 8, 13, 16, 24, 32, 48, 64, 125, 130]))
        padding = 'VALID' if (seed % 5 == 3 and not friendly and shape[1] > 4 and shape[2] > 4) else 'SAME'
        k = (rng.standard_normal((ksz, ksz, shape[3], oc)) / np.sqrt(ksz * ksz * shape[3])).astype(np.float32)
        spec = {"kernel": k, "padding": padding, "biases": (0.3 * rng.standard_normal(oc)).astype(np.float32)}
        if rng.random() < 0.7:
            gamma = rng.uniform(0.5, 2.0, oc).astype(np.float32)
            if rng.random() < 0.4:
                gamma *= rng.choice([-1.0, 1.0], oc).astype(np.float32)
            spec.update(mean=(0.2 * rng.standard_normal(oc)).astype(np.float32), variance=rng.uniform(0.3, 2.0, oc).astype(np.float32),
                        gamma=gamma, epsilon=1e-5, leaky=True)
        oh = shape[1] if padding == 'SAME' else shape[1] - ksz + 1
        ow = shape[2] if padding == 'SAME' else shape[2] - ksz + 1
        pool = [None, (2, 2, 'SAME'), (2, 1, 'SAME')][int(rng.integers(0, 3))]
        if pool is not None and min(oh, ow) < 2:
            pool = None
        spec["pool"] = pool
        if pool is not None and pool[1] == 2:
            oh, ow = (oh + 1) // 2, (ow + 1) // 2
        blocks.append(spec)
        shape = (batch, oh, ow, oc)
        if min(oh, ow) < 3:
            break
    x = rng.standard_normal((batch, h, w, ic)).astype(np.float32)
    return x, blocks


def _oracle(x, blocks):
    c = chain.Chain("oracle")
    cur = c.create_input(x.shape)
    for b in blocks:
        cur = c.create_conv2d(cur, b["kernel"], [1, 1, 1, 1], b["padding"])
        cur = c.create_bias_add(cur, b["biases"])
        if "gamma" in b:
            cur = c.create_batch_norm(cur, b["mean"], b["variance"], b["gamma"], 1e-5)
            cur = c.create_leaky_relu(cur)
        if b["pool"] is not None:
            k, s, pad = b["pool"]
            cur = c.create_max_pool2d(cur, [1, k, k, 1], [1, s, s, 1], pad)
    return c.run(x)


@pytest.mark.parametrize("seed", range(64))
def test_random_block_stacks_against_the_oracle(oracle_libs, seed):
    x, blocks = _random_case(seed)
    precision = ["tf32", "bf16"][(seed // 2) % 2]
    latency = seed % 3 == 0
    ref = _oracle(x, blocks)
    net = _lib.Net(x.shape, blocks, precision=precision, latency=latency)
    plan = [net.block_info(i)["kernel"] for i in range(net.num_blocks)]
    got = net.forward(x)
    again = net.forward(x)
    net.close()
    desc = "seed %d %s%s x%s %s plan=%s" % (seed, precision, " latency" if latency else "", x.shape,
                                            [(b["kernel"].shape, b["padding"], b["pool"], "gamma" in b) for b in blocks], plan)
    assert got.shape == ref.shape, desc
    np.testing.assert_array_equal(got, again, err_msg=desc)            # reproducible run to run, split-K included
    scale = max(float(np.abs(ref).max()), 1e-3)
    err = float(np.abs(got - ref).max()) / scale
    print("%s err=%.3g" % (desc, err))
    assert np.isfinite(got).all(), desc
    assert err < TOL[precision] * len(blocks), desc
