// conv0_umma.cuh -- the first conv block (3 input channels) on tcgen05:
//   3x3 SAME conv (3 -> 16) + bias + folded BatchNorm + LeakyReLU(0.1) + 2x2/2 max-pool,
// i.e. nodes conv2d_0 .. max_pool2d_0 of proj3/yolov2tiny.py:30-34, straight from the raw
// fp32 NHWC image (12-byte pixels, which TMA cannot address) to the pooled activation.
//
// No im2col tile is ever built.  A CTA keeps a PATCH of the input in shared memory, pixel-major
// with the channels padded to one 16-byte unit per pixel ({r,g,b,0} fp32, or {r,g,b,0,0,0,0,0}
// bf16).  In the UMMA "no-swizzle, K-major" layout a core matrix is 8 rows x 16 bytes with rows
// 16 bytes apart -- exactly 8 consecutive pixels of the patch -- and the two core matrices of
// one MMA's K slice may overlap (leading byte offset 16): row m of the A operand for the filter
// taps (dj, dj+1) is just patch bytes [16 (m+dj), 16 (m+dj) + 32).  So every tap is a descriptor
// offset into the same patch; the hardware does the gather.  (tools/probe_desc.cu, Q3.)
//
// Tile = one output row pair (2j, 2j+1) x 128 columns.  TMEM columns [0,16) accumulate output
// row 2j and [16,32) row 2j+1: each of the 4 patch rows r is multiplied once by the stacked
// weights [W(di=r) | W(di=r-1)] (N = 32; N = 16 for the first and last patch row).  The
// epilogue thread of lane m therefore holds both rows of pixel column m: the vertical max is
// thread-local, the horizontal max one shuffle.
//
// Warp roles (576 threads): 8 producer warps (each builds whole patches, tile t -> warp t % 8, so
// eight global-load latencies overlap: the producers are latency-bound), 2 MMA warps and 2 x 4
// epilogue warps (both alternate tiles).  A CTA owns a contiguous range of tiles, so every role walks
// (column tile, row pair, image) incrementally instead of dividing.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace conv0 {

// Ring discipline: every patch stage is always filled by the same producer warp and drained by the
// same MMA warp (kStages is a multiple of both warp counts), and every accumulator stage belongs to
// one MMA warp and one epilogue group.  Each waiter therefore follows its barrier phase by phase; a
// waiter that could get two phases ahead would see a stale parity as "done".
constexpr int kProducerWarps = 8;
constexpr int kMmaWarps = 2;       // tiles alternate between two issuing warps (small, issue-bound MMAs)
constexpr int kEpilogueGroups = 2;
constexpr int kThreads = (kProducerWarps + kMmaWarps + 4 * kEpilogueGroups) * 32;
constexpr int kStages = 8;         // patch ring
constexpr int kAccStages = 4;      // TMEM accumulator ring, 32 columns each
constexpr int kPatchPx = 136;      // pixels per patch row: [w0-4, w0+132)
constexpr int kGroups = kPatchPx / 4;
constexpr int kPatchRowBytes = kPatchPx * 16;
constexpr int kPatchBytes = 4 * kPatchRowBytes;
constexpr int kWBytes = 8 * 1024;  // 8 weight operands (4 patch rows x 2 tap pairs), <= 32 x 8 x 4 B each
constexpr int kSmemBytes = kStages * kPatchBytes + kWBytes + 1024 + 256;
static_assert(kStages % kProducerWarps == 0 && kStages % kMmaWarps == 0, "one owner warp per stage");
static_assert(kAccStages % kMmaWarps == 0 && kAccStages % kEpilogueGroups == 0, "one owner per accumulator stage");

struct Params {
    const float *in;     // [B, H, W, 3] fp32
    void *out;           // [B, H/2, W/2, 16]
    const void *wpack;   // 8 operands in canonical no-swizzle layout, 1 KB apart (see pack_weights)
    const float *scale;  // [16]
    const float *shift;  // [16]
    int B, H, W;
    int col_tiles;       // ceil(W / 128)
    int num_tiles;       // B * (H/2) * col_tiles
    int leaky, round_tf32;
    int pool_first;      // every scale[c] > 0: scale/shift, LeakyReLU and rounding are then monotone, so the
                         // epilogue pools the raw accumulators and activates one value instead of four
    float *prepool;      // debug: also dump the pre-pool activations [B,H,W,16] fp32
};

// no-swizzle K-major descriptor: LBO = byte distance between the two 16-byte K halves of one MMA,
// SBO = byte distance between 8-row groups
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           ((uint64_t)1 << 46);
}

template <bool BF16, bool OUT_BF16>
__global__ void __launch_bounds__(kThreads, 1) conv0_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint8_t *s_patch = smem;
    uint8_t *s_w = smem + kStages * kPatchBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_w + kWBytes);     // full[S], empty[S], tfull[A], tempty[A]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(s_w + kWBytes + 512);
    float *s_scale = reinterpret_cast<float *>(s_w + kWBytes + 1024);
    float *s_shift = s_scale + 16;

    const uint32_t bar_full = ptx::smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_tfull = bar_empty + 8 * kStages;
    const uint32_t bar_tempty = bar_tfull + 8 * kAccStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows2 = p.H >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { ptx::mbar_init(bar_full + 8 * s, 1); ptx::mbar_init(bar_empty + 8 * s, 1); }
        for (int s = 0; s < kAccStages; ++s) { ptx::mbar_init(bar_tfull + 8 * s, 1); ptx::mbar_init(bar_tempty + 8 * s, 128); }
        ptx::fence_barrier_init();
    }
    if (warp == kProducerWarps) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 32 * kAccStages);
    for (int i = threadIdx.x; i < kWBytes / 16; i += kThreads)
        reinterpret_cast<uint4 *>(s_w)[i] = reinterpret_cast<const uint4 *>(p.wpack)[i];
    if (threadIdx.x < 16) { s_scale[threadIdx.x] = p.scale[threadIdx.x]; s_shift[threadIdx.x] = p.shift[threadIdx.x]; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // weights were written by the generic proxy
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int per_cta = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tile_begin = (int)blockIdx.x * per_cta;
    const int my_tiles = max(0, min(p.num_tiles, tile_begin + per_cta) - tile_begin);
    // tile -> (column tile ct, row pair j, image n), advanced incrementally
    struct Coord { int ct, j, n; };
    auto decode = [&](int tile) {
        Coord c;
        c.ct = tile % p.col_tiles;
        c.j = (tile / p.col_tiles) % rows2;
        c.n = tile / (p.col_tiles * rows2);
        return c;
    };
    auto advance = [&](Coord &c, int steps) {
        c.ct += steps;
        while (c.ct >= p.col_tiles) {
            c.ct -= p.col_tiles;
            if (++c.j == rows2) { c.j = 0; ++c.n; }
        }
    };

    if (warp < kProducerWarps) {
        // ------------------------------------------------------------ patch producers
        // 136 groups of 4 pixels per patch; lane handles groups lane, lane+32, ... (5 rounds).  The
        // group -> (patch row, column group) split does not depend on the tile: hoist it.
        int g_row[5], g_src[5], g_dst[5], g_px[5];
#pragma unroll
        for (int it = 0; it < 5; ++it) {
            const int g = lane + 32 * it;
            const int r = g / kGroups, gx = g - r * kGroups;
            g_row[it] = g < 4 * kGroups ? r : 1 << 20;              // invalid groups fail the row test below
            g_px[it] = 4 * gx - 4;
            g_src[it] = (r * p.W + 4 * gx - 4) * 3;
            g_dst[it] = r * kPatchRowBytes + gx * 64;
        }
        Coord tc = decode(tile_begin + warp);
        for (int t = warp; t < my_tiles; t += kProducerWarps, advance(tc, kProducerWarps)) {
            const int j = tc.j, n = tc.n;
            const int w0 = tc.ct * 128;
            const int stage = t % kStages;
            const uint32_t phase = (t / kStages) & 1;
            // pointer to pixel (row 2j-1, column w0) of image n; only dereferenced where in range
            const float *base = p.in + ((size_t)n * p.H * p.W + (long long)(2 * j - 1) * p.W + w0) * 3;
            float4 v[5][3];
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const bool ok = (unsigned)(2 * j - 1 + g_row[it]) < (unsigned)p.H && (unsigned)(w0 + g_px[it]) < (unsigned)p.W;
                if (ok) {
                    const float4 *src = reinterpret_cast<const float4 *>(base + g_src[it]);
                    v[it][0] = __ldg(src); v[it][1] = __ldg(src + 1); v[it][2] = __ldg(src + 2);
                } else {
                    v[it][0] = v[it][1] = v[it][2] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);     // loads are already in flight
            uint8_t *patch = s_patch + stage * kPatchBytes;
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                if (g_row[it] < 4) {
                    uint4 *dst = reinterpret_cast<uint4 *>(patch + g_dst[it]);
                    const float f[12] = {v[it][0].x, v[it][0].y, v[it][0].z, v[it][0].w, v[it][1].x, v[it][1].y,
                                         v[it][1].z, v[it][1].w, v[it][2].x, v[it][2].y, v[it][2].z, v[it][2].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 o;
                        if constexpr (BF16) {
                            __nv_bfloat162 rg = __floats2bfloat162_rn(f[3 * q], f[3 * q + 1]);
                            __nv_bfloat162 b0 = __floats2bfloat162_rn(f[3 * q + 2], 0.f);
                            o.x = *reinterpret_cast<uint32_t *>(&rg);
                            o.y = *reinterpret_cast<uint32_t *>(&b0);
                            o.z = 0u; o.w = 0u;
                        } else {
                            o.x = __float_as_uint(ptx::round_tf32(f[3 * q]));
                            o.y = __float_as_uint(ptx::round_tf32(f[3 * q + 1]));
                            o.z = __float_as_uint(ptx::round_tf32(f[3 * q + 2]));
                            o.w = 0u;
                        }
                        dst[q] = o;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_full + 8 * stage);
        }
    } else if (warp < kProducerWarps + kMmaWarps) {
        // ------------------------------------------------------------ MMA issuers
        if (ptx::elect_one()) {
            constexpr uint32_t fmt = BF16 ? 1u : 2u;
            constexpr uint32_t idesc16 = ptx::umma_idesc(fmt, 16), idesc32 = ptx::umma_idesc(fmt, 32);
            // descriptor words: lo = start >> 4 | LBO >> 4 << 16, hi = SBO >> 4 | version; no swizzle
            constexpr uint32_t HI = (128u >> 4) | (1u << 14);
            const uint32_t w_lo0 = ptx::smem_u32(s_w) >> 4;
            for (int t = warp - kProducerWarps; t < my_tiles; t += kMmaWarps) {
                const int stage = t % kStages, acc = t % kAccStages;
                ptx::mbar_wait(bar_tempty + 8 * acc, ((t / kAccStages) & 1) ^ 1);
                ptx::mbar_wait(bar_full + 8 * stage, (t / kStages) & 1);
                ptx::tc_fence_after();
                const uint32_t a_lo0 = (ptx::smem_u32(s_patch + stage * kPatchBytes) >> 4) | ((16u >> 4) << 16);
                const uint32_t d = tmem_base + acc * 32;
                // patch rows 1 and 2 feed both output rows (N = 32), rows 0 and 3 one each (N = 16).
                // Operand q = r*2 + pair lives at s_w + q*1024; LBO: 32 rows x 16 B = 512 (256 for N = 16).
                // Row 1 goes first so its N = 32 MMA initialises all 32 accumulator columns.
#pragma unroll
                for (int oi = 0; oi < 4; ++oi) {
                    const int r = oi == 0 ? 1 : oi == 1 ? 2 : oi == 2 ? 0 : 3;
                    const bool wide = r == 1 || r == 2;
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr) {
                        // A rows: pixel m, taps dj = 2*pr, 2*pr+1  ->  patch pixel m + dj + 3
                        const uint32_t a_off = (r * kPatchRowBytes + (3 + 2 * pr) * 16) >> 4;
                        const uint32_t b_lo = w_lo0 + (((r * 2 + pr) * 1024) >> 4) + (((wide ? 512u : 256u) >> 4) << 16);
                        ptx::umma<BF16>(d + (r == 3 ? 16 : 0), ptx::umma_desc_join(a_lo0 + a_off, HI),
                                        ptx::umma_desc_join(b_lo, HI), wide ? idesc32 : idesc16, (oi | pr) != 0 ? 1u : 0u);
                    }
                }
                ptx::umma_commit(bar_empty + 8 * stage);
                ptx::umma_commit(bar_tfull + 8 * acc);
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (two groups alternate tiles)
        const int ew = warp - kProducerWarps - kMmaWarps;
        const int group = ew >> 2;
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
        const int m = quarter * 32 + lane;            // pixel column within the tile
        const int OW = p.W >> 1;
        const bool odd = lane & 1;
        float sc8[8], sh8[8];                          // scale/shift of the 8 channels this lane stores
#pragma unroll
        for (int c = 0; c < 8; ++c) { sc8[c] = s_scale[(odd ? 8 : 0) + c]; sh8[c] = s_shift[(odd ? 8 : 0) + c]; }
        Coord tc = decode(tile_begin + group);
        for (int t = group; t < my_tiles; t += kEpilogueGroups, advance(tc, kEpilogueGroups)) {
            const int ct = tc.ct, j = tc.j, n = tc.n;
            const int w = ct * 128 + m;
            const int acc = t % kAccStages;
            ptx::mbar_wait(bar_tfull + 8 * acc, (t / kAccStages) & 1);
            ptx::tc_fence_after();
            if (ct * 128 + quarter * 32 < p.W) {      // warp-uniform: skip quarters right of the image
                uint32_t raw[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 32, raw);
                ptx::tmem_ld_wait();
                float v[16];
                if (p.prepool != nullptr) {               // debug dump of the pooled-over activations (cold path)
                    for (int c = 0; c < 16; ++c) {
                        float a = fmaf(__uint_as_float(raw[c]), s_scale[c], s_shift[c]);
                        float b = fmaf(__uint_as_float(raw[16 + c]), s_scale[c], s_shift[c]);
                        if (p.leaky) { a = fmaxf(a, 0.1f * a); b = fmaxf(b, 0.1f * b); }
                        if constexpr (OUT_BF16) { a = __bfloat162float(__float2bfloat16_rn(a)); b = __bfloat162float(__float2bfloat16_rn(b)); }
                        else if (p.round_tf32) { a = ptx::round_tf32(a); b = ptx::round_tf32(b); }
                        if (w < p.W) {
                            p.prepool[(((size_t)n * p.H + 2 * j) * p.W + w) * 16 + c] = a;
                            p.prepool[(((size_t)n * p.H + 2 * j + 1) * p.W + w) * 16 + c] = b;
                        }
                    }
                }
                float r8[8];
                if (p.pool_first) {
                    // all scales positive: pool the raw accumulators (2 rows here, 2 columns by shuffle),
                    // then scale/shift + LeakyReLU + rounding on the 8 channels this lane stores
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float lo = fmaxf(__uint_as_float(raw[c]), __uint_as_float(raw[16 + c]));
                        const float hi = fmaxf(__uint_as_float(raw[8 + c]), __uint_as_float(raw[24 + c]));
                        const float keep = odd ? hi : lo;
                        const float send = odd ? lo : hi;
                        float y = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
                        y = fmaf(y, sc8[c], sh8[c]);
                        if (p.leaky) y = fmaxf(y, 0.1f * y);
                        if constexpr (OUT_BF16) y = __bfloat162float(__float2bfloat16_rn(y));
                        else if (p.round_tf32) y = ptx::round_tf32(y);
                        r8[c] = y;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        float a = fmaf(__uint_as_float(raw[c]), s_scale[c], s_shift[c]);
                        float b = fmaf(__uint_as_float(raw[16 + c]), s_scale[c], s_shift[c]);
                        if (p.leaky) { a = fmaxf(a, 0.1f * a); b = fmaxf(b, 0.1f * b); }
                        float mx = fmaxf(a, b);                      // vertical max: rows 2j, 2j+1
                        if constexpr (OUT_BF16) mx = __bfloat162float(__float2bfloat16_rn(mx));
                        else if (p.round_tf32) mx = ptx::round_tf32(mx);
                        v[c] = mx;                                   // rounding is monotone: max then round == round then max
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float keep = odd ? v[8 + c] : v[c];
                        const float send = odd ? v[c] : v[8 + c];
                        r8[c] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));   // horizontal max: columns w, w^1
                    }
                }
                if (w < p.W) {
                    const size_t pix = ((size_t)n * rows2 + j) * OW + (w >> 1);
                    if constexpr (OUT_BF16) {
                        __nv_bfloat16 *o = static_cast<__nv_bfloat16 *>(p.out) + pix * 16 + (odd ? 8 : 0);
                        __nv_bfloat162 t0 = __floats2bfloat162_rn(r8[0], r8[1]), t1 = __floats2bfloat162_rn(r8[2], r8[3]);
                        __nv_bfloat162 t2 = __floats2bfloat162_rn(r8[4], r8[5]), t3 = __floats2bfloat162_rn(r8[6], r8[7]);
                        uint4 pk;
                        pk.x = *reinterpret_cast<uint32_t *>(&t0); pk.y = *reinterpret_cast<uint32_t *>(&t1);
                        pk.z = *reinterpret_cast<uint32_t *>(&t2); pk.w = *reinterpret_cast<uint32_t *>(&t3);
                        *reinterpret_cast<uint4 *>(o) = pk;
                    } else {
                        float *o = static_cast<float *>(p.out) + pix * 16 + (odd ? 8 : 0);
                        *reinterpret_cast<float4 *>(o) = make_float4(r8[0], r8[1], r8[2], r8[3]);
                        *reinterpret_cast<float4 *>(o + 4) = make_float4(r8[4], r8[5], r8[6], r8[7]);
                    }
                }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(bar_tempty + 8 * acc);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kProducerWarps) ptx::tmem_dealloc(tmem_base, 32 * kAccStages);
}

// Host side: the 8 weight operands (patch row r = 0..3, tap pair pr = 0..1), each in the canonical
// no-swizzle K-major layout  elem(n, k) -> (k / KC) * LBO + (n / 8) * 128 + (n % 8) * 16 + (k % KC) * ES
// with KC = elements per 16 bytes, LBO = 16 * rows.  Rows n < 16 hold W[di = r] (output row 2j),
// rows 16..31 hold W[di = r - 1] (output row 2j+1); patch row 0 has only the first, row 3 only the
// second half.  K index = (dj - 2*pr) * KC + c, channels c >= 3 and tap dj = 3 are zero.
// kernel: HWIO [3][3][3][16] fp32.  rnd(): value -> operand bit pattern (tf32-rounded fp32 or bf16).
template <typename Elem, typename Rnd>
void pack_weights(const float *kernel, Elem *out /* kWBytes */, Rnd rnd) {
    static_assert(sizeof(Elem) == 4 || sizeof(Elem) == 2, "fp32 (tf32) or bf16 operands");
    for (int i = 0; i < kWBytes / (int)sizeof(Elem); ++i) out[i] = rnd(0.f);
    for (int r = 0; r < 4; ++r)
        for (int pr = 0; pr < 2; ++pr) {
            const bool wide = r == 1 || r == 2;
            const int rows = wide ? 32 : 16;
            const int lbo = 16 * rows;
            Elem *op = out + (size_t)(r * 2 + pr) * 1024 / sizeof(Elem);
            for (int nrow = 0; nrow < rows; ++nrow) {
                int di, oc;
                if (r == 0) { di = 0; oc = nrow; }
                else if (r == 3) { di = 2; oc = nrow; }
                else if (nrow < 16) { di = r; oc = nrow; }
                else { di = r - 1; oc = nrow - 16; }
                for (int half = 0; half < 2; ++half) {
                    const int dj = 2 * pr + half;
                    if (dj > 2) continue;
                    for (int c = 0; c < 3; ++c) {
                        const float wv = kernel[((size_t)(di * 3 + dj) * 3 + c) * 16 + oc];
                        const size_t byte = (size_t)half * lbo + (nrow / 8) * 128 + (nrow % 8) * 16 + c * sizeof(Elem);
                        op[byte / sizeof(Elem)] = rnd(wv);
                    }
                }
            }
        }
}

}  // namespace conv0
