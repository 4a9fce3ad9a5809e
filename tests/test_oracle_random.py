"""CPU property test: the C restatement (oracle/oracle.c) against the reference's own C compiled unmodified
(oracle/_ref, proj3/dnn_openblas.c) on seeded random shapes the YOLO stack never visits -- odd sizes, 1x1 /
3x3 / 5x5 kernels, strides 1 and 2, SAME and VALID padding, windows that straddle the padded border.
Batch 1 only: the reference is wrong beyond that.  Element-wise ops and the pool
must agree bit for bit, the convolution to OpenBLAS' summation order."""
import numpy as np
import pytest

from conftest import normalised_error
import chain

CASES = list(range(24))


def _both(x, fn):
    outs = []
    for backend in ("oracle", "ref_openblas"):
        c = chain.Chain(backend)
        node = fn(c, c.create_input(x.shape))
        c.set_out_node(node)
        outs.append(c.run(x))
    return outs


@pytest.mark.parametrize("case", CASES)
def test_random_conv_matches_compiled_reference(oracle_libs, case):
    if not chain.available("ref_openblas"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    rng = np.random.default_rng(1000 + case)
    k = int(rng.choice([1, 3, 5]))
    s = int(rng.choice([1, 2]))
    padding = str(rng.choice(["SAME", "VALID"]))
    h, w = (int(v) for v in rng.integers(k, 23, size=2))
    ic, oc = int(rng.integers(1, 10)), int(rng.integers(1, 13))
    x = rng.standard_normal((1, h, w, ic)).astype(np.float32)
    kern = rng.standard_normal((k, k, ic, oc)).astype(np.float32)
    got, ref = _both(x, lambda c, i: c.create_conv2d(i, kern, [1, s, s, 1], padding))
    assert got.shape == ref.shape, (got.shape, ref.shape, k, s, padding)
    assert normalised_error(got, ref) < 2e-5, (k, s, padding, h, w, ic, oc)


@pytest.mark.parametrize("case", CASES)
def test_random_pool_and_elementwise_are_bit_exact(oracle_libs, case):
    if not chain.available("ref_openblas"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    rng = np.random.default_rng(2000 + case)
    k = int(rng.choice([2, 3]))
    s = int(rng.choice([1, 2]))
    padding = str(rng.choice(["SAME", "VALID"]))
    h, w = (int(v) for v in rng.integers(k, 19, size=2))
    c_ = int(rng.integers(1, 20))
    x = rng.standard_normal((1, h, w, c_)).astype(np.float32)
    x[0, rng.integers(0, h), rng.integers(0, w)] = x[0, 0, 0]           # exact ties for the first-max rule
    bias = rng.standard_normal(c_).astype(np.float32)
    mean = rng.standard_normal(c_).astype(np.float32)
    var = rng.uniform(0.1, 2.0, c_).astype(np.float32)
    gamma = rng.standard_normal(c_).astype(np.float32)                    # negative gammas included

    def chain_of_ops(c, i):
        n = c.create_bias_add(i, bias)
        n = c.create_batch_norm(n, mean, var, gamma, 1e-5)
        n = c.create_leaky_relu(n)
        return c.create_max_pool2d(n, [1, k, k, 1], [1, s, s, 1], padding)

    got, ref = _both(x, chain_of_ops)
    np.testing.assert_array_equal(got, ref, err_msg=str((k, s, padding, h, w, c_)))
