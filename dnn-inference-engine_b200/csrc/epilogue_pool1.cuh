// epilogue_pool1.cuh -- the 2x2 stride-1 SAME max-pool of block 5 (max_pool2d_5, proj3/yolov2tiny.py:64:
// ksize 2, stride 1 on 13x13 -> pad 0 top/left, 1 bottom/right with finfo(float32).min,
// proj3/dnn_openblas.py:232-235) fused into the epilogue of the im2col kernels.
//
// An im2col tile is 128 CONSECUTIVE output pixels m = (n*H + h)*W + w, so the window of pixel m is rows
// m, m+1, m+W, m+W+1 of the same tile as long as m+W+1 is inside it: tiles advance by 128 - (W+1) rows
// (`m_stride`) and the last W+1 rows of every tile are halo, recomputed by the next tile.  Per 32-column
// chunk the four epilogue warps park their activated rows in a [128][36]-float staging area (16-byte
// stores, conflict-free; reads with lane = channel are conflict-free too), then warp q emits output rows
// 32q .. 32q+31: it loads its 32 + W + 1 staged rows once, all loads in flight together, takes the
// horizontal pair maxima p[r] = max(s[r], s[r+1]) and the vertical ones max(p[r], p[r+W]) in registers, and
// stores one 128-byte line per row.  Neighbours outside the image are skipped, which is what the reference's
// -FLT_MAX padding amounts to; max() does not depend on the (di, dj) order the reference visits
// (dnn_openblas.c:209-222).  The image width is a compile-time constant (13: the YOLOv2-tiny grid) so that
// the whole window walk unrolls; the planner keeps the separate pool kernel for any other width.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace pool1 {

constexpr int kW = 13;                           // image width the fused form is compiled for
constexpr int kHalo = kW + 1;
constexpr int kPitch = 36;                       // floats per staged row (144 bytes: 16-byte aligned rows)
constexpr int kStageBytes = 128 * kPitch * 4;

__device__ __forceinline__ float lds(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// v: columns [col0, col0 + 32) of tile row 32 * quarter + lane, activated and rounded to the storage type.
// m0: output pixel of tile row 0.  Rows [0, m_stride) of the tile are stored; all 128 epilogue threads call
// this together (two named-barrier syncs).  prepool (debug): also dump the pooled-over activations [B,H,W,OC] fp32.
template <bool OUT_BF16>
__device__ __forceinline__ void store_chunk(uint32_t s_stage, const float (&v)[32], int quarter, int lane, int m0,
                                            int m_stride, int M_total, int H, int OC, int col0, void *out,
                                            float *prepool, size_t out_elems) {
    const uint32_t mine = s_stage + (uint32_t)(quarter * 32 + lane) * (kPitch * 4);
#pragma unroll
    for (int j = 0; j < 32; j += 4) sts4(mine + 4 * j, v[j], v[j + 1], v[j + 2], v[j + 3]);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int r0 = quarter * 32;
    // staged rows r0 .. r0 + 32 + kW + 1 of this lane's channel (rows past the tile are never used: clamp the address)
    constexpr int kRows = 32 + kHalo + 1;
    float x[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) x[i] = lds(s_stage + (uint32_t)(min(r0 + i, 127) * kPitch + lane) * 4);
    asm volatile("bar.sync 1, 128;" ::: "memory");     // everybody has its rows: the staging area may be rewritten
    const int m = m0 + r0;
    const int hw = m % (H * kW);
    int h = hw / kW, w = hw - h * kW;
    // horizontal pair maxima; row i and row i + kW are the same image column, so one walk over w serves both
    int wi = w;
#pragma unroll
    for (int i = 0; i < 32 + kW; ++i) {
        if (wi + 1 < kW) x[i] = fmaxf(x[i], x[i + 1]);
        if (++wi == kW) wi = 0;
    }
    const int col = col0 + lane;
    if (col >= OC) return;
    const int n_rows = min(min(32, m_stride - r0), M_total - m);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i < n_rows) {
            // the debug dump wants the un-pooled value: it was overwritten above only when it had a right neighbour
            const float a = (h + 1 < H) ? fmaxf(x[i], x[i + kW]) : x[i];
            if constexpr (OUT_BF16) *ptx::out_at<__nv_bfloat16>(out, (size_t)(m + i) * OC + col, 1, out_elems, 14) = __float2bfloat16_rn(a);
            else *ptx::out_at<float>(out, (size_t)(m + i) * OC + col, 1, out_elems, 15) = a;
        }
        if (++w == kW) { w = 0; if (++h == H) h = 0; }
    }
    (void)prepool;
}

// BF16 storage: the same walk on packed pairs.  v0 / v1: columns [col0, col0 + 32) and [col0 + 32, col0 + 64) of the
// thread's row, activated (fp32).  They are rounded to bf16 here -- the reference pools the stored values -- and staged
// as 32 words of two channels each, so that a lane owns two adjacent channels: half the staging traffic, half the
// window-walk instructions (HMNMX2) and 4-byte instead of 2-byte stores per lane (a whole 128-byte line per row and
// warp) for the same 64 channels.  With 2-byte stores the BF16 block 5 was epilogue-bound (tensor pipe 49 % busy).
__device__ __forceinline__ void store_chunk_bf16x2(uint32_t s_stage, const float (&v0)[32], const float (&v1)[32], int quarter,
                                                   int lane, int m0, int m_stride, int M_total, int H, int OC, int col0,
                                                   void *out, size_t out_elems) {
    const uint32_t mine = s_stage + (uint32_t)(quarter * 32 + lane) * (kPitch * 4);
    auto pack = [](float a, float b) {
        const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<const uint32_t *>(&t);
    };
#pragma unroll
    for (int j = 0; j < 16; j += 4)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(mine + 4 * j), "r"(pack(v0[2 * j], v0[2 * j + 1])),
                     "r"(pack(v0[2 * j + 2], v0[2 * j + 3])), "r"(pack(v0[2 * j + 4], v0[2 * j + 5])),
                     "r"(pack(v0[2 * j + 6], v0[2 * j + 7])) : "memory");
#pragma unroll
    for (int j = 0; j < 16; j += 4)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(mine + 64 + 4 * j), "r"(pack(v1[2 * j], v1[2 * j + 1])),
                     "r"(pack(v1[2 * j + 2], v1[2 * j + 3])), "r"(pack(v1[2 * j + 4], v1[2 * j + 5])),
                     "r"(pack(v1[2 * j + 6], v1[2 * j + 7])) : "memory");
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int r0 = quarter * 32;
    constexpr int kRows = 32 + kHalo + 1;
    __nv_bfloat162 x[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
        uint32_t w;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(s_stage + (uint32_t)(min(r0 + i, 127) * kPitch + lane) * 4));
        x[i] = *reinterpret_cast<const __nv_bfloat162 *>(&w);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");     // everybody has its rows: the staging area may be rewritten
    const int m = m0 + r0;
    const int hw = m % (H * kW);
    int h = hw / kW, w = hw - h * kW;
    int wi = w;
#pragma unroll
    for (int i = 0; i < 32 + kW; ++i) {
        if (wi + 1 < kW) x[i] = __hmax2(x[i], x[i + 1]);
        if (++wi == kW) wi = 0;
    }
    const int col = col0 + 2 * lane;
    if (col >= OC) return;
    const int n_rows = min(min(32, m_stride - r0), M_total - m);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i < n_rows) {
            const __nv_bfloat162 a = (h + 1 < H) ? __hmax2(x[i], x[i + kW]) : x[i];
            *reinterpret_cast<__nv_bfloat162 *>(ptx::out_at<__nv_bfloat16>(out, (size_t)(m + i) * OC + col, 2, out_elems, 18)) = a;
        }
        if (++w == kW) { w = 0; if (++h == H) h = 0; }
    }
}

// The 2x2 stride-2 pool over an im2col tile that is whole PAIRS of image rows (m_stride = 2k W pixels, W even;
// small batches under the latency plan, where spatial tiles would be mostly padding): pooled pixel (pr, pc) of the
// tile is the maximum of tile rows 2 pr W + 2 pc + {0, 1, W, W + 1}.  `s_stage` holds [128][kPitch] activated values
// of `ncols` <= 32 channels starting at channel col0; the caller has synchronised after filling it.  Image heights
// are even, so a pair of flattened rows never straddles two images and the pooled tensor [B, H/2, W/2, OC] is simply
// flattened-row / 2.
template <bool OUT_BF16>
__device__ __forceinline__ void emit_s2(uint32_t s_stage, int quarter, int lane, int m0, int m_stride, int M_total, int W,
                                        int OC, int col0, int ncols, void *out, size_t out_elems) {
    const int half_w = W >> 1;
    const int n_pooled = m_stride >> 2;               // <= 32: at most 8 per warp
    const int col = col0 + lane;
    if (lane >= ncols || col >= OC) return;
    // pooled flattened row of the tile's first row pair; everything below is tile-local 32-bit arithmetic and the
    // (row, column) of window q = quarter + 4k advances by counting -- at batch 1 one warp per scheduler runs this and
    // eight integer divisions were most of the phase (470 instructions, 2.5 us)
    const size_t base = (size_t)((m0 / W) >> 1) * half_w * OC + col;
    // all windows of this warp first (up to 8 x 4 loads in flight), then the maxima and the stores: one window after
    // the other, each behind its four dependent loads, made this the longest phase of the batch-1 kernels
    float v[8][4];
    uint32_t o[8];
    bool live[8];
    int pr = 0, pc = quarter;
    while (pc >= half_w) { pc -= half_w; ++pr; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int q = quarter + 4 * k;
        const int r00 = 2 * pr * W + 2 * pc;
        live[k] = q < n_pooled && m0 + r00 + W + 1 < M_total;       // windows past the batch are skipped
        o[k] = (uint32_t)(pr * half_w + pc) * (uint32_t)OC;
        if (live[k]) {
            const uint32_t a = s_stage + (uint32_t)(r00 * kPitch + lane) * 4;
            v[k][0] = lds(a);
            v[k][1] = lds(a + kPitch * 4);
            v[k][2] = lds(a + W * kPitch * 4);
            v[k][3] = lds(a + (W + 1) * kPitch * 4);
        }
        pc += 4;
        while (pc >= half_w) { pc -= half_w; ++pr; }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (live[k]) {
            const float r = fmaxf(fmaxf(v[k][0], v[k][1]), fmaxf(v[k][2], v[k][3]));
            if constexpr (OUT_BF16) *ptx::out_at<__nv_bfloat16>(out, base + o[k], 1, out_elems, 16) = __float2bfloat16_rn(r);
            else *ptx::out_at<float>(out, base + o[k], 1, out_elems, 17) = r;
        }
    }
}

// The stride-1 window walk on an already filled staging area (`ncols` <= 32 channels): second half of store_chunk.
template <bool OUT_BF16>
__device__ __forceinline__ void emit_s1(uint32_t s_stage, int quarter, int lane, int m0, int m_stride, int M_total, int H,
                                        int OC, int col0, int ncols, void *out, size_t out_elems) {
    const int r0 = quarter * 32;
    constexpr int kRows = 32 + kHalo + 1;
    float x[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) x[i] = lds(s_stage + (uint32_t)(min(r0 + i, 127) * kPitch + lane) * 4);
    const int m = m0 + r0;
    const int hw = m % (H * kW);
    int h = hw / kW, w = hw - h * kW;
    int wi = w;
#pragma unroll
    for (int i = 0; i < 32 + kW; ++i) {
        if (wi + 1 < kW) x[i] = fmaxf(x[i], x[i + 1]);
        if (++wi == kW) wi = 0;
    }
    const int col = col0 + lane;
    if (lane >= ncols || col >= OC) return;
    const int n_rows = min(min(32, m_stride - r0), M_total - m);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i < n_rows) {
            const float a = (h + 1 < H) ? fmaxf(x[i], x[i + kW]) : x[i];
            if constexpr (OUT_BF16) *ptx::out_at<__nv_bfloat16>(out, (size_t)(m + i) * OC + col, 1, out_elems, 14) = __float2bfloat16_rn(a);
            else *ptx::out_at<float>(out, (size_t)(m + i) * OC + col, 1, out_elems, 15) = a;
        }
        if (++w == kW) { w = 0; if (++h == H) h = 0; }
    }
}

// One 32-column chunk through the staging area with either pool (stride 1: store_chunk above; stride 2 here).
template <bool OUT_BF16>
__device__ __forceinline__ void store_chunk_s2(uint32_t s_stage, const float (&v)[32], int quarter, int lane, int m0, int m_stride,
                                               int M_total, int W, int OC, int col0, void *out, size_t out_elems) {
    const uint32_t mine = s_stage + (uint32_t)(quarter * 32 + lane) * (kPitch * 4);
#pragma unroll
    for (int j = 0; j < 32; j += 4) sts4(mine + 4 * j, v[j], v[j + 1], v[j + 2], v[j + 3]);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    emit_s2<OUT_BF16>(s_stage, quarter, lane, m0, m_stride, M_total, W, OC, col0, 32, out, out_elems);
    asm volatile("bar.sync 1, 128;" ::: "memory");
}

// Debug variant of the first half: dumps this thread's own row of the chunk (all 128 tile rows are written by
// somebody: rows [0, m_stride) belong to this tile, the halo rows to the next one, which writes the same values).
__device__ __forceinline__ void dump_prepool(const float (&v)[32], int m, int M_total, int owned, int OC, int col0,
                                             float *prepool) {
    if (prepool == nullptr || !owned || m >= M_total) return;
    for (int j = 0; j < 32; ++j)
        if (col0 + j < OC) prepool[(size_t)m * OC + col0 + j] = v[j];
}

}  // namespace pool1
