// postproc.cuh -- the detection-head decoding that follows the layer stack in the reference
// (`postprocessing`, proj3/yolov2tiny.py:94-180; `iou` :182-200; `non_maximal_suppression`
// :202-226; `sigmoid` :228; `softmax` :231), batched on the device: one CTA per image turns the
// [13,13,125] head output into the final, NMS-filtered label boxes, so that only a few hundred
// bytes per image travel back to the host instead of the 84.5 KB head tensor.
//
// Arithmetic follows the reference step by step in float32 (NumPy >= 2 keeps its scalar math in
// float32; np.e is rounded to float32 before the power), with explicit round-to-nearest
// intrinsics so that nvcc cannot contract multiplies and adds into FMAs.  exp and pow are
// evaluated in double and rounded once, i.e. correctly rounded float32 results; numpy's own
// float32 exp is 1-2 ulp off that, so scores agree to a few ulp and every discrete decision
// (class, truncated box edge, threshold, order, IoU test) agrees wherever it is not within an ulp
// of flipping.  The class-probability sum uses numpy's pairwise order for 20 elements
// (8 lanes, tree, 4-element tail).  Box edges saturate at +-2^29 so that the IoU integer
// arithmetic (the reference uses unbounded Python ints) is exact in int64.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace postproc {

constexpr int kGrid = 13;
constexpr int kBoxes = 5;
constexpr int kClasses = 20;
constexpr int kFields = 5 + kClasses;
constexpr int kCand = kGrid * kGrid * kBoxes;     // 845
constexpr int kThreads = 256;
constexpr float kCoordLimit = 536870912.0f;       // 2^29

struct Det {                // == dnnb_det
    int cls, left, top, right, bottom;
    float score;
};

struct Params {
    const float *pred;      // [B][13][13][125]
    const Det *cand_in;     // optional: [B][845] candidates to use instead of decoding `pred`
    Det *cand_out;          // optional: [B][845] dump of the decoded candidates (before the threshold)
    Det *dets;              // [B][max_det]
    int *counts;            // [B] boxes kept (may exceed max_det; only the first max_det are stored)
    int B;
    int max_det;
    float score_thr;        // compared in float32, as numpy compares a float32 with a Python float
    double iou_thr;         // compared in double, as Python compares the true-division result
};

// ln(float32(e)): np.e ** y in float32 is powf(2.71828174591064453125f, y)
__device__ __constant__ double kLogE32 = 0.99999996963214;
__device__ __constant__ float kAnchors[2 * kBoxes] = {1.08f, 1.19f, 3.42f, 4.41f, 6.63f, 11.38f, 9.42f, 5.11f, 16.62f, 10.52f};

__device__ __forceinline__ float exp_rn(float y) { return (float)exp((double)y); }
__device__ __forceinline__ float sigmoid_ref(float x) {          // 1 / (1 + np.e ** -x)
    const float p = (float)exp((double)(-x) * kLogE32);
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, p));
}
__device__ __forceinline__ int trunc_edge(float v) {             // int(v), saturating
    if (!(v == v)) return 0;
    if (v >= kCoordLimit) return (int)kCoordLimit;
    if (v <= -kCoordLimit) return -(int)kCoordLimit;
    return (int)v;
}

__device__ __forceinline__ Det decode_candidate(const float *cell, int row, int col, int b) {
    const float tx = cell[0], ty = cell[1], tw = cell[2], th = cell[3], tc = cell[4];
    const float cx = __fmul_rn(__fadd_rn((float)col, sigmoid_ref(tx)), 32.0f);
    const float cy = __fmul_rn(__fadd_rn((float)row, sigmoid_ref(ty)), 32.0f);
    const float w = __fmul_rn(__fmul_rn(exp_rn(tw), kAnchors[2 * b]), 32.0f);
    const float h = __fmul_rn(__fmul_rn(exp_rn(th), kAnchors[2 * b + 1]), 32.0f);
    const float conf = sigmoid_ref(tc);

    float m = cell[5];
#pragma unroll
    for (int i = 1; i < kClasses; ++i) m = fmaxf(m, cell[5 + i]);
    float e[kClasses];
#pragma unroll
    for (int i = 0; i < kClasses; ++i) e[i] = exp_rn(__fsub_rn(cell[5 + i], m));
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(e[j], e[8 + j]);
    float sum = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
#pragma unroll
    for (int i = 16; i < kClasses; ++i) sum = __fadd_rn(sum, e[i]);
    int best = 0;
    float best_p = __fdiv_rn(e[0], sum);
#pragma unroll
    for (int i = 1; i < kClasses; ++i) {
        const float pi = __fdiv_rn(e[i], sum);
        if (pi > best_p) { best_p = pi; best = i; }      // strict: the first maximum wins
    }
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);     // roi / 2. is exact
    Det d;
    d.cls = best;
    d.left = trunc_edge(__fsub_rn(cx, hw));
    d.top = trunc_edge(__fsub_rn(cy, hh));
    d.right = trunc_edge(__fadd_rn(cx, hw));
    d.bottom = trunc_edge(__fadd_rn(cy, hh));
    d.score = __fmul_rn(conf, best_p);
    return d;
}

// IoU > threshold with the reference's conventions: +1 pixel extents, an empty intersection
// is NOT clamped, Python int arithmetic (exact here: |edge| <= 2^29 keeps every product
// below 2^62), int -> float conversions and one double division.
__device__ __forceinline__ bool iou_exceeds(const Det &a, const Det &b, double thr) {
    const long long xa = max(a.left, b.left), ya = max(a.top, b.top);
    const long long xb = min(a.right, b.right), yb = min(a.bottom, b.bottom);
    const long long inter = (xb - xa + 1) * (yb - ya + 1);
    const long long area_a = ((long long)a.right - a.left + 1) * ((long long)a.bottom - a.top + 1);
    const long long area_b = ((long long)b.right - b.left + 1) * ((long long)b.bottom - b.top + 1);
    const long long uni = area_a + area_b - inter;
    // a float quotient settles every case that is not within 1e-3 of the threshold; the rest (and NaN)
    // take the reference's arithmetic: int -> double conversions and one double division.
    // uni == 0 raises ZeroDivisionError in the reference; here IEEE division makes it +-inf / NaN
    const float q = __fdividef((float)inter, (float)uni), tf = (float)thr;
    if (fabsf(q - tf) > 1e-3f * (fabsf(tf) + 1.0f) && fabsf(q) < 1e30f) return q > tf;
    return __ddiv_rn((double)inter, (double)uni) > thr;
}

// One CTA per image (grid-stride over the batch).
__global__ void __launch_bounds__(kThreads) detect_kernel(const Params p) {
    __shared__ Det s_cand[kCand];          // compacted survivors of the score threshold, raster order
    __shared__ unsigned short s_sorted[kCand];
    __shared__ unsigned short s_kept[kCand];
    __shared__ int s_warp_count[kThreads / 32];
    __shared__ int s_n;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int img = blockIdx.x; img < p.B; img += gridDim.x) {
        // ---- decode + threshold + order-preserving compaction, 256 candidates per pass
        if (tid == 0) s_n = 0;
        __syncthreads();
        for (int base = 0; base < kCand; base += kThreads) {
            const int c = base + tid;
            Det d{};
            bool pass = false;
            if (c < kCand) {
                if (p.cand_in != nullptr) {
                    d = p.cand_in[(size_t)img * kCand + c];
                } else {
                    const int cellidx = c / kBoxes, b = c - cellidx * kBoxes;
                    const int row = cellidx / kGrid, col = cellidx - row * kGrid;
                    float cell[kFields];
                    const float *src = p.pred + ((size_t)img * kGrid * kGrid + cellidx) * (kBoxes * kFields) + b * kFields;
#pragma unroll
                    for (int i = 0; i < kFields; ++i) cell[i] = __ldg(src + i);
                    d = decode_candidate(cell, row, col, b);
                }
                if (p.cand_out != nullptr) p.cand_out[(size_t)img * kCand + c] = d;
                pass = d.score > p.score_thr;
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, pass);
            if (lane == 0) s_warp_count[warp] = __popc(ballot);
            __syncthreads();
            int offset = s_n;
            for (int w = 0; w < warp; ++w) offset += s_warp_count[w];
            if (pass) s_cand[offset + __popc(ballot & ((1u << lane) - 1))] = d;
            __syncthreads();
            if (tid == 0) {
                int total = 0;
                for (int w = 0; w < kThreads / 32; ++w) total += s_warp_count[w];
                s_n += total;
            }
            __syncthreads();
        }
        const int n = s_n;

        // ---- stable sort by score, descending: rank = how many candidates come before this one
        for (int i = tid; i < n; i += kThreads) {
            const float si = s_cand[i].score;
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const float sj = s_cand[j].score;
                rank += (sj > si) || (sj == si && j < i);
            }
            s_sorted[rank] = (unsigned short)i;
        }
        __syncthreads();

        // ---- greedy NMS, sequential over the sorted candidates, kept boxes checked 32 at a time
        if (warp == 0) {
            int nk = 0;
            for (int t = 0; t < n; ++t) {
                const Det cur = s_cand[s_sorted[t]];
                bool drop = false;
                for (int k0 = 0; k0 < nk && !drop; k0 += 32) {       // any overlap decides: stop at the first chunk that has one
                    const int k = k0 + lane;
                    const bool hit = k < nk && iou_exceeds(cur, s_cand[s_kept[k]], p.iou_thr);
                    drop = __any_sync(0xffffffffu, hit);
                }
                if (!drop) {
                    if (lane == 0) {
                        s_kept[nk] = s_sorted[t];
                        if (nk < p.max_det) p.dets[(size_t)img * p.max_det + nk] = cur;
                    }
                    ++nk;
                }
                __syncwarp();
            }
            if (lane == 0) p.counts[img] = nk;
            // rows past the kept boxes are defined too (zeros): the caller's buffer never sees stale device memory
            for (int k = nk + lane; k < p.max_det; k += 32) p.dets[(size_t)img * p.max_det + k] = Det{};
        }
        __syncthreads();
    }
}

}  // namespace postproc
