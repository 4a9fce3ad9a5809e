// conv_rowpair.cuh -- 3x3 SAME conv + bias/BatchNorm/LeakyReLU + 2x2/2 max-pool for the wide,
// few-channel blocks (blocks 1 and 2 of proj3/yolov2tiny.py:36-46: 208x208x16->32, 104x104x32->64),
// which are bound by memory traffic, not by the tensor pipe.
//
// conv_umma.cuh fetches one A tile per filter tap, i.e. it pulls every input pixel through L2
// nine times.  Here the CTA loads a PATCH (4 input rows x 130 pixels, all channels: one TMA box,
// zero-filled outside the image = SAME padding) once per tile and all nine taps are UMMA
// descriptors that start at different rows of that patch: a swizzled K-major operand may start
// at any row of a TMA-written slab (tools/probe_desc.cu, Q1/Q2).  The 3x3xICxOC weights stay
// resident in shared memory for the whole kernel.
//
// Tile = TR output rows x 128 columns (TR = 2, 4 or 8: as many as shared memory and TMEM allow;
// taller tiles read each patch row from shared memory fewer times per output pixel and amortise
// the per-tile handshakes).  TMEM columns [OC o, OC o + OC) hold output row o of the tile; patch
// row r feeds the output rows o = r - di, so it is multiplied once by the stacked weights
// [W(di=2) | W(di=1) | W(di=0)] (N = 3 OC, clipped at the tile's top and bottom) into the columns
// starting at OC max(0, r-2).  The epilogue thread of lane m holds every row of pixel column m:
// vertical max in registers, horizontal max by one shuffle.
//
// CTA = 608 threads: warp 0 TMA producer, warps 1-2 MMA issuers (alternate tiles; warp 1 owns TMEM),
// 4 x 4 epilogue warps; group g takes tiles g, g + 4, ... (block 1 was epilogue-bound with two groups: each warp
// is a chain of dependent TMEM loads, maxima, shuffles and stores that keeps its scheduler busy 70 % of the time).  One channel chunk only: IC * elem_size must equal the swizzle span (32/64/128 B).
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda.h>
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace rowpair {

constexpr int kMmaWarps = 2;      // tiles alternate between two issuing warps: the MMAs are small, issue-bound (four warps,
                                  // measured in round 2 on blocks with four stages and slots: 16->32 TF32 0.44 -> 0.49 ms,
                                  // BF16 32->64 0.158 -> 0.197 ms -- as in the first-block kernel, more issuers are slower)
constexpr int kEpiGroups = 4;     // epilogue groups of four warps (one per TMEM lane quarter); group g takes tiles g, g + 4, ...
constexpr int kThreads = (1 + kMmaWarps + 4 * kEpiGroups) * 32;
constexpr int kPatchW = 130;      // 128 columns + left/right halo

// LINEAR tiling (template flag): instead of 128-column tiles of the image (416, 208 and 104 columns are
// 3.25, 1.63 and 0.81 such tiles: 19 % of the MMA rows are padding) a tile is 128 consecutive positions
// of the sequence  (image, row group of TR rows, column position p in [0, W+4))  -- input columns
// -3 .. W (zero outside the image), output column p - 2.  Plane s of the patch (input row TR j + s - 1 of
// every row group j) is then one TMA *im2col* walk of 130 positions that wraps from one row group
// to the next (traversal stride TR in H, bounding box [s-1, H-1+s-TR]), taps dj = 0..2 are still row
// offsets into the plane, and a lane's rows of the tile are still the same column of consecutive
// image rows, so the pooling does not change.  4 positions per row group are padding: 2-4 %.
struct LinMaps {
    CUtensorMap m[10];    // one per patch plane (TR + 2 <= 10)
};

struct Params {
    int B, H, W;          // conv output = input spatial size
    int col_tiles;        // ceil(W / 128)
    int num_tiles;        // B * (H/TR) * col_tiles; LINEAR: ceil(B * (H/TR) * (W+4) / 128)
    const float *scale;   // [OC]
    const float *shift;   // [OC]
    void *out;            // [B, H/2, W/2, OC]
    int leaky, round_tf32;
    int pool_first;       // every scale[c] > 0: pool the raw accumulators, then activate one value instead of four
    float *prepool;       // debug: also dump the pre-pool activations [B,H,W,OC] fp32
    size_t out_elems;     // extent of `out` in elements (checked by -DDNNB_BOUNDS builds)
};

constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int pow2_at_least(int v) { int p = 32; while (p < v) p *= 2; return p; }

template <int OC, int SWZ, int TR, bool LINEAR = false>
struct Cfg {
    static constexpr int PATCH_ROWS = TR + 2;                  // output rows + top/bottom halo
    // pixels between patch planes: 130; LINEAR loads every plane with its own TMA request, whose
    // destination must be 128-byte aligned (32-byte pixels: 132)
    static constexpr int PITCH_PX = (LINEAR && (kPatchW * SWZ) % 128 != 0) ? kPatchW + 2 : kPatchW;
    static constexpr int PATCH_ROW = PITCH_PX * SWZ;
    static constexpr int PATCH_RAW = PATCH_ROWS * kPatchW * SWZ;   // bytes the TMA loads of one tile deliver
    static constexpr int PATCH_BYTES = (PATCH_ROWS * PATCH_ROW + 1023) / 1024 * 1024;
    static constexpr int W_DJ_BYTES = 3 * OC * SWZ;            // one dj: rows [di=2 | di=1 | di=0] x OC
    static constexpr int W_BYTES = 3 * W_DJ_BYTES;
    static constexpr int TAIL_BYTES = 1024 + 2 * OC * 4;
    static constexpr int BUDGET = 227 * 1024 - 1024;
    static constexpr int STAGES = cmin(4, (BUDGET - W_BYTES - TAIL_BYTES) / PATCH_BYTES) / kMmaWarps * kMmaWarps;
    static constexpr int SMEM_BYTES = STAGES * PATCH_BYTES + W_BYTES + TAIL_BYTES + 1024;
    static constexpr int ACC_COLS = TR * OC;
    static constexpr int ACC_STAGES = cmin(4, 512 / ACC_COLS);
    static_assert(TR % 2 == 0 && 3 * OC <= 256, "whole pooling row pairs; N = 3*OC must be one MMA");
    static constexpr int TMEM_COLS = pow2_at_least(ACC_COLS * ACC_STAGES);
    static_assert(STAGES >= 2, "need two patch stages");
    // one owner per ring slot (see conv0_umma.cuh): stages and accumulators never change their MMA warp
    static_assert(STAGES % kMmaWarps == 0 && ACC_STAGES % kMmaWarps == 0, "one owner warp per ring slot");
    // epilogue groups that share one accumulator slot (each takes a share of the tile's pooling row pairs): a group
    // visits every tile of its slot, so the slot's barrier never gets two phases ahead of a waiter
    static constexpr int GROUPS_PER_ACC = kEpiGroups > ACC_STAGES ? kEpiGroups / ACC_STAGES : 1;
    static_assert(ACC_STAGES % kEpiGroups == 0 || kEpiGroups % ACC_STAGES == 0, "one epilogue group per accumulator slot");
    static_assert((TR / 2) % GROUPS_PER_ACC == 0, "row pairs divide among the groups of a slot");
    static_assert(ACC_STAGES >= 2 && TMEM_COLS <= 512, "need two accumulator stages");
    static_assert(W_DJ_BYTES % 1024 == 0, "weight slabs must keep the swizzle phase");
    static_assert(OC % 16 == 0 && 3 * OC <= 256, "weights of one dj must fit one TMA box");
};

template <bool BF16, bool OUT_BF16, int OC, int SWZ, int TR, bool LINEAR = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_rowpair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                    const Params p, const __grid_constant__ LinMaps lin) {
    using C = Cfg<OC, SWZ, TR, LINEAR>;
    constexpr int STAGES = C::STAGES, ACC = C::ACC_STAGES;
    constexpr int KC = SWZ / 32;                     // MMAs along K per tap (32 bytes of K each)

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *s_w = smem + STAGES * C::PATCH_BYTES;
    uint8_t *tail = s_w + C::W_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail);     // full[S], empty[S], tfull[A], tempty[A], wbar
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tail + 512);
    float *s_scale = reinterpret_cast<float *>(tail + 1024);
    float *s_shift = s_scale + OC;

    const uint32_t patch_base = ptx::smem_u32(smem);
    const uint32_t w_base = ptx::smem_u32(s_w);
    const uint32_t bar_full = ptx::smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * STAGES;
    const uint32_t bar_tempty = bar_tfull + 8 * ACC;
    const uint32_t bar_w = bar_tempty + 8 * ACC;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows2 = p.H / TR;          // row tiles per image
    // a CTA owns a contiguous range of tiles; (column tile, row pair, image) advance incrementally
    const int per_cta = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tile_begin = (int)blockIdx.x * per_cta;
    const int my_tiles = max(0, min(p.num_tiles, tile_begin + per_cta) - tile_begin);
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

    const int P = p.W + 4;                // LINEAR: positions per row group
    ptx::griddep_launch();
    if (threadIdx.x == 0) {
        if constexpr (LINEAR) {
            for (int s = 0; s < C::PATCH_ROWS; ++s) ptx::prefetch_tensormap(&lin.m[s]);
        } else {
            ptx::prefetch_tensormap(&map_a);
        }
        ptx::prefetch_tensormap(&map_w);
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(bar_full + 8 * s, 1); ptx::mbar_init(bar_empty + 8 * s, 1); }
        for (int s = 0; s < ACC; ++s) { ptx::mbar_init(bar_tfull + 8 * s, 1); ptx::mbar_init(bar_tempty + 8 * s, 128 * C::GROUPS_PER_ACC); }
        ptx::mbar_init(bar_w, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), C::TMEM_COLS);
    for (int i = threadIdx.x; i < OC; i += kThreads) { s_scale[i] = p.scale[i]; s_shift[i] = p.shift[i]; }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp != 0) ptx::griddep_wait();

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(bar_w, C::W_BYTES);                 // resident weights, once
            for (int dj = 0; dj < 3; ++dj) ptx::tma_load_3d(w_base + dj * C::W_DJ_BYTES, &map_w, bar_w, 0, 0, dj);
            ptx::griddep_wait();                                    // the activations are the preceding kernel's output
            Coord tc = decode(tile_begin);
            for (int t = 0; t < my_tiles; ++t, advance(tc, 1)) {
                const int stage = t % STAGES;
                ptx::mbar_wait(bar_empty + 8 * stage, ((t / STAGES) & 1) ^ 1);
                ptx::mbar_expect_tx(bar_full + 8 * stage, C::PATCH_RAW);
                if constexpr (LINEAR) {
                    const int q0 = (tile_begin + t) * 128;                // first position of the tile (< 2^31: checked by the planner)
                    const int rg = q0 / P, pos = q0 - rg * P;
                    const int n = rg / rows2, j = rg - n * rows2;
#pragma unroll
                    for (int s = 0; s < C::PATCH_ROWS; ++s)
                        ptx::tma_load_im2col_4d(patch_base + stage * C::PATCH_BYTES + s * C::PATCH_ROW, &lin.m[s],
                                                bar_full + 8 * stage, 0, pos - 3, TR * j + s - 1, n, 0, 0);
                } else {
                    ptx::tma_load_4d(patch_base + stage * C::PATCH_BYTES, &map_a, bar_full + 8 * stage, 0, tc.ct * 128 - 1,
                                     TR * tc.j - 1, tc.n);
                }
            }
        }
    } else if (warp <= kMmaWarps) {
        // ------------------------------------------------------------ MMA issuers
        if (ptx::elect_one()) {
            constexpr uint32_t fmt = BF16 ? 1u : 2u;
            constexpr uint32_t HI = ptx::umma_desc_hi(SWZ);
            const uint32_t b_lo0 = ptx::umma_desc_lo(w_base);
            ptx::mbar_wait(bar_w, 0);
            for (int t = warp - 1; t < my_tiles; t += kMmaWarps) {
                const int stage = t % STAGES, acc = t % ACC;
                ptx::mbar_wait(bar_tempty + 8 * acc, ((t / ACC) & 1) ^ 1);
                ptx::mbar_wait(bar_full + 8 * stage, (t / STAGES) & 1);
                ptx::tc_fence_after();
                const uint32_t a_lo0 = ptx::umma_desc_lo(patch_base + stage * C::PATCH_BYTES);
                const uint32_t d = tmem_base + acc * C::ACC_COLS;
                // Weight rows of one dj slab: [W(di=2) | W(di=1) | W(di=0)], OC rows each.  Patch row r feeds
                // output rows o in [max(0, r-2), min(TR-1, r)] with di = r - o: operand rows from
                // OC (2 - (r - o_lo)), N = OC (#rows), accumulator columns from OC o_lo.
                // Rows 2, 5, 8, ... go first: together they touch every column exactly once and initialise it.
                // Every descriptor is (per-tile base) + (compile-time constant), in 16-byte units.
                constexpr int NINIT = (TR + 2) / 3;                  // rows 2, 5, ... <= TR + 1
#pragma unroll
                for (int oi = 0; oi < C::PATCH_ROWS; ++oi) {
                    int r;
                    if (oi < NINIT) r = 2 + 3 * oi;
                    else { const int q = oi - NINIT; r = q + q / 2; }   // the remaining rows 0,1,3,4,6,7,9,...
                    const int o_lo = r - 2 > 0 ? r - 2 : 0;
                    const int o_hi = r < TR - 1 ? r : TR - 1;
                    const int n = OC * (o_hi - o_lo + 1);
                    const int wrow = OC * (2 - (r - o_lo));
                    const uint32_t idesc = ptx::umma_idesc(fmt, n);
                    const uint32_t dcol = d + OC * o_lo;
#pragma unroll
                    for (int dj = 0; dj < 3; ++dj) {
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            const uint32_t a_off = ((r * C::PITCH_PX + dj) * SWZ + 32 * k) >> 4;
                            const uint32_t b_off = (dj * C::W_DJ_BYTES + wrow * SWZ + 32 * k) >> 4;
                            ptx::umma<BF16>(dcol, ptx::umma_desc_join(a_lo0 + a_off, HI), ptx::umma_desc_join(b_lo0 + b_off, HI),
                                            idesc, (oi >= NINIT || dj != 0 || k != 0) ? 1u : 0u);
                        }
                    }
                }
                ptx::umma_commit(bar_empty + 8 * stage);
                ptx::umma_commit(bar_tfull + 8 * acc);
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (two groups alternate tiles)
        const int group = (warp - 1 - kMmaWarps) >> 2;
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const int OW = p.W >> 1;
        const bool odd = lane & 1;
        // Three epilogue variants as three separate tile loops, chosen once (inside one loop the compiler if-converts
        // the variants and every thread issues all of them: ~920 issue slots per 16-channel step for ~100 useful ones):
        //   0: every scale > 0 -> pool the raw accumulators, then activate (the normal case)
        //   1: activate, then pool (negative scales)          2: as 1, and dump the pooled-over activations (debug)
        auto epilogue_tiles = [&](auto MODE) {
            constexpr int mode = decltype(MODE)::value;
            // group g serves accumulator slot g % SLOTS, i.e. tiles g % SLOTS, + SLOTS, ...; with fewer slots than groups
            // GROUPS_PER_ACC groups share a tile, each taking RP_PER of its pooling row pairs
            constexpr int SLOTS = kEpiGroups / C::GROUPS_PER_ACC;            // tiles in flight over the groups
            constexpr int RP_PER = TR / 2 / C::GROUPS_PER_ACC;
            const int rp_begin = (group / SLOTS) * RP_PER;
            Coord tc = decode(tile_begin + group % SLOTS);
            for (int t = group % SLOTS; t < my_tiles; t += SLOTS, advance(tc, SLOTS)) {
                int ct = tc.ct, j = tc.j, n = tc.n;
                int w = ct * 128 + m;
                bool live = ct * 128 + quarter * 32 < p.W;     // warp-uniform: skip quarters right of the image
                if constexpr (LINEAR) {
                    const int q = (tile_begin + t) * 128 + m;
                    const int rg = q / P;
                    n = rg / rows2;
                    j = rg - n * rows2;
                    w = q - rg * P - 2;                        // output column; the 4 pad positions of a row group fail 0 <= w < W
                    if (n >= p.B) w = -1;
                    live = true;
                }
                const int acc = t % ACC;
                ptx::mbar_wait(bar_tfull + 8 * acc, (t / ACC) & 1);
                ptx::tc_fence_after();
                if (live) {
#pragma unroll 1
                    for (int rp = rp_begin; rp < rp_begin + RP_PER; ++rp) {     // this group's pooling row pairs of the tile
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * C::ACC_COLS + 2 * rp * OC;
                    const int orow = TR * j + 2 * rp;                      // conv output row of the pair's first row
                    const size_t pix = ((size_t)n * (p.H >> 1) + (orow >> 1)) * OW + (w >> 1);
#pragma unroll 1
                    for (int c0 = 0; c0 < OC; c0 += 16) {
                        uint32_t top[16], bot[16];
                        ptx::tmem_ld_32x16(taddr + c0, top);         // first row of the pair
                        ptx::tmem_ld_32x16(taddr + OC + c0, bot);    // second row
                        ptx::tmem_ld_wait();
                        float v[16];
                        if (mode == 2 && (unsigned)w < (unsigned)p.W) {     // debug dump of the pooled-over activations (cold path)
                            for (int c = 0; c < 16; ++c) {
                                const float sc = s_scale[c0 + c], sh = s_shift[c0 + c];
                                float a = fmaf(__uint_as_float(top[c]), sc, sh);
                                float b = fmaf(__uint_as_float(bot[c]), sc, sh);
                                if (p.leaky) { a = fmaxf(a, 0.1f * a); b = fmaxf(b, 0.1f * b); }
                                if constexpr (OUT_BF16) { a = __bfloat162float(__float2bfloat16_rn(a)); b = __bfloat162float(__float2bfloat16_rn(b)); }
                                else if (p.round_tf32) { a = ptx::round_tf32(a); b = ptx::round_tf32(b); }
                                p.prepool[(((size_t)n * p.H + orow) * p.W + w) * OC + c0 + c] = a;
                                p.prepool[(((size_t)n * p.H + orow + 1) * p.W + w) * OC + c0 + c] = b;
                            }
                        }
                        float r8[8];
                        if constexpr (mode == 0) {
                            // all scales positive: max over the 2x2 window of raw accumulators first
                            const float4 sca = *reinterpret_cast<const float4 *>(s_scale + c0 + (odd ? 8 : 0));
                            const float4 scb = *reinterpret_cast<const float4 *>(s_scale + c0 + (odd ? 8 : 0) + 4);
                            const float4 sha = *reinterpret_cast<const float4 *>(s_shift + c0 + (odd ? 8 : 0));
                            const float4 shb = *reinterpret_cast<const float4 *>(s_shift + c0 + (odd ? 8 : 0) + 4);
                            const float sc8[8] = {sca.x, sca.y, sca.z, sca.w, scb.x, scb.y, scb.z, scb.w};
                            const float sh8[8] = {sha.x, sha.y, sha.z, sha.w, shb.x, shb.y, shb.z, shb.w};
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float lo = fmaxf(__uint_as_float(top[c]), __uint_as_float(bot[c]));
                                const float hi = fmaxf(__uint_as_float(top[8 + c]), __uint_as_float(bot[8 + c]));
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
                                const float sc = s_scale[c0 + c], sh = s_shift[c0 + c];
                                float a = fmaf(__uint_as_float(top[c]), sc, sh);
                                float b = fmaf(__uint_as_float(bot[c]), sc, sh);
                                if (p.leaky) { a = fmaxf(a, 0.1f * a); b = fmaxf(b, 0.1f * b); }
                                float mx = fmaxf(a, b);                  // vertical max; rounding is monotone, so round after
                                if constexpr (OUT_BF16) mx = __bfloat162float(__float2bfloat16_rn(mx));
                                else if (p.round_tf32) mx = ptx::round_tf32(mx);
                                v[c] = mx;
                            }
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float keep = odd ? v[8 + c] : v[c];
                                const float send = odd ? v[c] : v[8 + c];
                                r8[c] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
                            }
                        }
                        if ((unsigned)w < (unsigned)p.W) {
                            const size_t off = pix * OC + c0 + (odd ? 8 : 0);
                            if constexpr (OUT_BF16) {
                                __nv_bfloat162 t0 = __floats2bfloat162_rn(r8[0], r8[1]), t1 = __floats2bfloat162_rn(r8[2], r8[3]);
                                __nv_bfloat162 t2 = __floats2bfloat162_rn(r8[4], r8[5]), t3 = __floats2bfloat162_rn(r8[6], r8[7]);
                                uint4 pk;
                                pk.x = *reinterpret_cast<uint32_t *>(&t0); pk.y = *reinterpret_cast<uint32_t *>(&t1);
                                pk.z = *reinterpret_cast<uint32_t *>(&t2); pk.w = *reinterpret_cast<uint32_t *>(&t3);
                                *reinterpret_cast<uint4 *>(ptx::out_at<__nv_bfloat16>(p.out, off, 8, p.out_elems, 5)) = pk;
                            } else {
                                float *o = ptx::out_at<float>(p.out, off, 8, p.out_elems, 6);
                                *reinterpret_cast<float4 *>(o) = make_float4(r8[0], r8[1], r8[2], r8[3]);
                                *reinterpret_cast<float4 *>(o + 4) = make_float4(r8[4], r8[5], r8[6], r8[7]);
                            }
                        }
                    }
                    }   // rp
                }
                ptx::tc_fence_before();
                ptx::mbar_arrive(bar_tempty + 8 * acc);
            }
        };
        if (p.prepool != nullptr) epilogue_tiles(std::integral_constant<int, 2>{});
        else if (p.pool_first) epilogue_tiles(std::integral_constant<int, 0>{});
        else epilogue_tiles(std::integral_constant<int, 1>{});
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace rowpair
