/* oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's proj3 CPU algorithm for the
 * YOLOv2-tiny hot path (Conv2D, BiasAdd, BatchNorm, LeakyReLU(0.1), MaxPool2D,
 * NHWC float32).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (dnn-inference-engine_b200/) never does and has no CPU fallback.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/cs492-projects/).  The restatement differs from the reference
 * only where the reference is broken for batch > 1 (SURVEY.md Appendix B, items
 * B1-B3): here every op indexes images correctly and never writes to its inputs.
 *
 * Pinning: tests/test_oracle.py checks this file against tests/golden/*.npz,
 * which were produced by importing the reference's own proj3/dnn_openblas.py
 * over its own dnn_openblas.c compiled unmodified (oracle/make_golden.py).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define ORACLE_MAX(x, y) ((x) >= (y) ? (x) : (y)) /* proj3/dnn_openblas.c:7 -- accumulator first, ties keep it */

/* TensorFlow SAME/VALID geometry.  proj3/dnn_openblas.py:127-142 (get_out_pads). */
void oracle_out_pads(int in_size, int filter, int stride, int same,
                     int *out_size, int *pad_front, int *pad_back)
{
    if (same) {
        int o = (in_size + stride - 1) / stride;
        int pad = (o - 1) * stride + filter - in_size;
        if (pad < 0) pad = 0;
        *out_size = o;
        *pad_front = pad / 2;
        *pad_back = pad - pad / 2;
    } else {
        *out_size = (in_size - filter + 1 + stride - 1) / stride;
        *pad_front = 0;
        *pad_back = 0;
    }
}

/* np.pad(x, [(0,0),(pt,pb),(pl,pr),(0,0)], 'constant', value) as the host layer does
 * before conv (proj3/dnn_openblas.py:176-179, value 0) and before pooling
 * (proj3/dnn_openblas.py:232-235, value finfo(float32).min). */
void oracle_pad_nhwc(const float *in, float *out, int batch, int h, int w, int c,
                     int pt, int pb, int pl, int pr, float value)
{
    int ph = h + pt + pb, pw = w + pl + pr;
    size_t total = (size_t)batch * ph * pw * c;
    for (size_t i = 0; i < total; ++i) out[i] = value;
    for (int b = 0; b < batch; ++b)
        for (int i = 0; i < h; ++i)
            memcpy(out + (((size_t)b * ph + i + pt) * pw + pl) * c,
                   in + ((size_t)b * h + i) * w * c, sizeof(float) * (size_t)w * c);
}

/* proj3/dnn_openblas.c:135-158 (im2col): one row of `col` per output pixel, the
 * K index runs channel-major: k = c*(kh*kw) + di*kw + dj. `im` is one padded image. */
void oracle_im2col(const float *im, float *col, int oh, int ow, int ih, int iw, int ic,
                   int kh, int kw, int sh, int sw)
{
    int K = ic * kh * kw;
    (void)ih;
    for (int i = 0; i < oh; ++i)
        for (int j = 0; j < ow; ++j) {
            float *row = col + (size_t)(i * ow + j) * K;
            for (int c = 0; c < ic; ++c)
                for (int k = 0; k < kh * kw; ++k)
                    row[c * kh * kw + k] =
                        im[((size_t)(i * sh + k / kw) * iw + (j * sw + k % kw)) * ic + c];
        }
}

/* C[M x N] = A[M x K] * B[K x N], row-major, float32 accumulate: the operation the
 * reference asks of cblas_sgemm at proj3/dnn_openblas.c:184-192 (alpha 1, beta 0).
 * OpenBLAS 's summation order is not specified by the reference; this loop adds in
 * increasing k, which agrees with it to ~1e-6 relative (tests/test_oracle.py). */
void oracle_sgemm(const float *A, const float *B, float *C, int M, int N, int K)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M; ++i) {
        float *c = C + (size_t)i * N;
        for (int j = 0; j < N; ++j) c[j] = 0.0f;
        for (int k = 0; k < K; ++k) {
            float a = A[(size_t)i * K + k];
            const float *b = B + (size_t)k * N;
            for (int j = 0; j < N; ++j) c[j] += a * b[j];
        }
    }
}

/* proj3/dnn_openblas.c:160-194 (conv2d_mul): per image im2col then GEMM with
 * kernel_r = kernel(HWIO).transpose(2,0,1,3).reshape(K, od) (proj3/dnn_openblas.py:166-167).
 * `in` is already padded to [batch, ih, iw, ic]; `col` is scratch [oh*ow, ic*kh*kw].
 * Unlike the reference (:170) the input pointer advances by ih*iw*ic per image. */
void oracle_conv2d_mul(const float *in, float *col, const float *kernel_r, float *result,
                       int batch, int oh, int ow, int od, int ih, int iw, int ic,
                       int kh, int kw, int sh, int sw)
{
    for (int b = 0; b < batch; ++b) {
        oracle_im2col(in + (size_t)b * ih * iw * ic, col, oh, ow, ih, iw, ic, kh, kw, sh, sw);
        oracle_sgemm(col, kernel_r, result + (size_t)b * oh * ow * od, oh * ow, od, ic * kh * kw);
    }
}

/* proj3/dnn_openblas.c:9-38 (bias_add): copy, then result[..., d] += bias[d] * 1.0f. */
void oracle_bias_add(const float *in, const float *biases, float *result,
                     int batch, int h, int w, int d)
{
    size_t pixels = (size_t)batch * h * w;
    for (size_t p = 0; p < pixels; ++p)
        for (int k = 0; k < d; ++k)
            result[p * d + k] = in[p * d + k] + biases[k] * 1.0f;
}

/* proj3/dnn_openblas.c:40-65 (batch_norm): ((x - mean) / sqrt(var + eps)) * gamma, where
 * the square root is taken in double and stored back as float (:48-50).  The reference
 * overwrites its variance argument and drops the batch offset (:56-58); this does neither. */
void oracle_batch_norm(const float *in, const float *mean, const float *variance,
                       const float *gamma, float epsilon, float *result,
                       int batch, int oh, int ow, int od)
{
    size_t pixels = (size_t)batch * oh * ow;
    for (size_t p = 0; p < pixels; ++p)
        for (int d = 0; d < od; ++d) {
            float sd = (float)sqrt((double)(variance[d] + epsilon));
            result[p * od + d] = ((in[p * od + d] - mean[d]) / sd) * gamma[d];
        }
}

/* proj3/dnn_openblas.c:236-254 (leaky_relu): t < 0 ? 0.1 * t : t with a double literal. */
void oracle_leaky_relu(const float *in, float *result, int batch, int oh, int ow, int od)
{
    size_t n = (size_t)batch * oh * ow * od;
    for (size_t i = 0; i < n; ++i) {
        float t = in[i];
        result[i] = t < 0 ? (float)(0.1 * (double)t) : t;
    }
}

/* proj3/dnn_openblas.c:196-234 (max_pool2d) over an input already padded with
 * finfo(float32).min; first maximum in (di, dj) raster order wins (MAX macro :7).
 * `argmax`, when not NULL, receives that raster index di*kw+dj -- the "max-pool index"
 * BASELINE.json asks to be bit-identical. */
void oracle_max_pool2d(const float *in, float *result, int *argmax,
                       int batch, int oh, int ow, int od, int ih, int iw, int ic,
                       int kh, int kw, int sh, int sw)
{
    for (int b = 0; b < batch; ++b)
        for (int i = 0; i < oh; ++i)
            for (int j = 0; j < ow; ++j)
                for (int d = 0; d < od; ++d) {
                    const float *base = in + (((size_t)b * ih + i * sh) * iw + j * sw) * ic + d;
                    float imax = base[0];
                    int where = 0;
                    for (int di = 0; di < kh; ++di)
                        for (int dj = 0; dj < kw; ++dj) {
                            float v = base[((size_t)di * iw + dj) * ic];
                            if (!(imax >= v)) { imax = v; where = di * kw + dj; }
                        }
                    size_t ri = (((size_t)b * oh + i) * ow + j) * od + d;
                    result[ri] = imax;
                    if (argmax) argmax[ri] = where;
                }
}
