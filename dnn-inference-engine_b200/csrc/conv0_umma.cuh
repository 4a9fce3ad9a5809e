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
// Tile = 8 output rows x 128 columns of one image (patch: 10 input rows x 136 pixels).  TMEM
// columns [16 o, 16 o + 16) accumulate output row o of the tile.  Patch row r feeds the output
// rows o = r - di, so it is multiplied ONCE by the stacked weights [W(di=2) | W(di=1) | W(di=0)]
// (N = 48, clipped to 32 / 16 at the tile's top and bottom) into columns starting at 16 max(0, r-2).
// An MMA's K slice is two 16-byte units = two (patch row, tap) pairs, and a patch row has three taps: the slice
// [dj 0 | dj 1] of row r is one MMA (leading byte offset 16: the next pixel), and the left-over taps dj 2 of rows r
// and r + 1 share another one (leading byte offset = one patch row) against a 64-row operand that covers the four
// output rows r-2 .. r+1 the two of them feed -- 15 MMAs per 1024 output pixels instead of 20 with a half-empty
// [dj 2 | -] slice per row.  What bounds this kernel is the shared-memory data path (the MMAs' operand fetch is
// more than half of it, profiles/r02_SUMMARY.md), so a quarter fewer operand bytes is a quarter less of that.  The epilogue thread of lane m holds all 8 rows of pixel column
// m: vertical max in registers, horizontal max by one shuffle with lane m ^ 1.
//
// The tiles are small, so every per-tile handshake counts: measured on B200 the barrier skeleton
// alone costs ~200 cycles per tile, which is why a tile is 8 rows and not one pooling row pair.
//
// Warp roles (608 threads): one loader thread streams the raw 12-byte pixels global -> shared with bulk
// async copies (one per patch row, kRawDepth tiles ahead); 8 producer warps convert them into the padded
// patch; 2 MMA warps and 2 x 4 epilogue warps alternate tiles.  The loads must not be issued by the
// converting threads: their generic->async proxy fence (a MEMBAR) would wait for every load in flight
// and serialise the prefetch -- measured: one global-load latency per tile.  A CTA owns a contiguous range of tiles and walks
// (column tile, row tile, image) incrementally.
//
// Ring discipline: every patch stage and accumulator stage has the same producer(s) and the same
// consumer warp for the whole kernel, each of which visits it phase by phase; a waiter that could
// get two mbarrier phases ahead would read a stale parity as "complete".
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace conv0 {

constexpr int kProducerWarps = 8;
constexpr int kMmaWarps = 2;
constexpr int kEpilogueGroups = 2;
constexpr int kLoaderWarp = kProducerWarps + kMmaWarps + 4 * kEpilogueGroups;   // one thread of it issues the bulk copies
constexpr int kThreads = (kLoaderWarp + 1) * 32;
constexpr int kTileRows = 8;                       // output rows per tile
constexpr int kPatchRows = kTileRows + 2;
constexpr int kStages = 4;                         // patch ring
constexpr int kAccStages = 4;                      // TMEM accumulator ring, 16 * kTileRows columns each
constexpr int kAccCols = 16 * kTileRows;
constexpr int kPatchPx = 136;                      // pixels per patch row: [w0-4, w0+132)
constexpr int kGroups = kPatchPx / 4;              // 4-pixel groups per patch row
constexpr int kTileGroups = kGroups * kPatchRows;  // 340
constexpr int kRounds = (kTileGroups + kProducerWarps * 32 - 1) / (kProducerWarps * 32);   // 2
constexpr int kPatchRowBytes = kPatchPx * 16;
constexpr int kPatchBytes = kPatchRows * kPatchRowBytes;
constexpr int kWOperandBytes = 2 * 48 * 16;        // operand 0, taps (dj 0, dj 1) of one patch row: 2 K halves x 48 rows x 16 B
constexpr int kW2OperandBytes = 2 * 64 * 16;       // operand 1, tap dj 2 of TWO consecutive patch rows: 2 K halves x 64 rows
constexpr int kWBytes = kWOperandBytes + kW2OperandBytes;
constexpr int kRawDepth = 8;                       // raw-pixel staging ring: tiles whose global loads are in flight
constexpr int kRawRowBytes = kPatchPx * 12;        // one patch row of raw 12-byte pixels (= kGroups * 48)
constexpr int kRawBytes = (kPatchRows * kRawRowBytes + 127) / 128 * 128;
// uint8 frames (IN_U8): a raw patch row is 160 3-byte pixels [w0-16, w0+144) = 480 bytes, so that both
// ends of every bulk copy fall on 16-byte boundaries (needs W % 16 == 0)
constexpr int kRawPxU8 = 160;
constexpr int kRawLeadU8 = 16;
constexpr int kRawRowBytesU8 = kRawPxU8 * 3;
static_assert(kPatchRows * kRawRowBytesU8 <= kRawBytes, "uint8 rows fit the raw slots");
constexpr int kSmemBytes = kStages * kPatchBytes + kRawDepth * kRawBytes + kWBytes + 1024 + 256 + 1024 /*uint8 LUT*/;
static_assert(kStages % kMmaWarps == 0 && kAccStages % kMmaWarps == 0 && kAccStages % kEpilogueGroups == 0,
              "one owner warp per ring slot");
static_assert(kAccCols * kAccStages <= 512, "accumulators must fit TMEM");

struct Params {
    const void *in;      // [B, H, W, 3] fp32, or uint8 frames (IN_U8)
    const uint32_t *lut; // IN_U8: operand bit pattern of float32(v / 255.) for v = 0..255, already rounded
                         // the way the fp32 path rounds (tf32: + half ulp; bf16: RN, in the low 16 bits)
    int swap_rb;         // IN_U8: frames are BGR, the network wants RGB (proj3/__init__.py:11)
    void *out;           // [B, H/2, W/2, 16]
    const void *wpack;   // 2 stacked operands in canonical no-swizzle layout (see pack_weights)
    const float *scale;  // [16]
    const float *shift;  // [16]
    int B, H, W;         // H % 8 == 0, W % 4 == 0
    int col_tiles;       // column tiles per row tile: ceil(W / 128)
    int num_tiles;       // B * (H / 8) * col_tiles
    int leaky, round_tf32;
    int pool_first;      // every scale[c] > 0: scale/shift, LeakyReLU and rounding are then monotone, so the
                         // epilogue pools the raw accumulators and activates one value instead of four
    float *prepool;      // debug: also dump the pre-pool activations [B,H,W,16] fp32
    size_t out_elems;    // extent of `out` in elements (checked by -DDNNB_BOUNDS builds)
};

template <bool BF16, bool OUT_BF16, bool IN_U8 = false>
__global__ void __launch_bounds__(kThreads, 1) conv0_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint8_t *s_patch = smem;
    uint8_t *s_raw = smem + kStages * kPatchBytes;
    uint8_t *s_w = s_raw + kRawDepth * kRawBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_w + kWBytes);     // full[S], empty[S], tfull[A], tempty[A], rfull[D], rempty[D]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(s_w + kWBytes + 512);
    float *s_scale = reinterpret_cast<float *>(s_w + kWBytes + 1024);
    float *s_shift = s_scale + 16;
    uint32_t *s_lut = reinterpret_cast<uint32_t *>(s_w + kWBytes + 1024 + 256);

    const uint32_t bar_full = ptx::smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_tfull = bar_empty + 8 * kStages;
    const uint32_t bar_tempty = bar_tfull + 8 * kAccStages;
    const uint32_t bar_rfull = bar_tempty + 8 * kAccStages;
    const uint32_t bar_rempty = bar_rfull + 8 * kRawDepth;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row_tiles = p.H / kTileRows;

    ptx::griddep_launch();
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(bar_full + 8 * s, kProducerWarps);
            ptx::mbar_init(bar_empty + 8 * s, 1);
        }
        for (int s = 0; s < kAccStages; ++s) { ptx::mbar_init(bar_tfull + 8 * s, 1); ptx::mbar_init(bar_tempty + 8 * s, 128); }
        for (int s = 0; s < kRawDepth; ++s) { ptx::mbar_init(bar_rfull + 8 * s, 1); ptx::mbar_init(bar_rempty + 8 * s, kProducerWarps); }
        ptx::fence_barrier_init();
    }
    if (warp == kProducerWarps) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), kAccCols * kAccStages);
    for (int i = threadIdx.x; i < kWBytes / 16; i += kThreads)
        reinterpret_cast<uint4 *>(s_w)[i] = reinterpret_cast<const uint4 *>(p.wpack)[i];
    if (threadIdx.x < 16) { s_scale[threadIdx.x] = p.scale[threadIdx.x]; s_shift[threadIdx.x] = p.shift[threadIdx.x]; }
    if constexpr (IN_U8) {
        if (threadIdx.x < 256) s_lut[threadIdx.x] = p.lut[threadIdx.x];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // weights were written by the generic proxy
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::griddep_wait();      // weights, constants and table are loaded; frames and the output buffer are stream-ordered

    const int per_cta = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tile_begin = (int)blockIdx.x * per_cta;
    const int my_tiles = max(0, min(p.num_tiles, tile_begin + per_cta) - tile_begin);
    // tile -> (column tile ct, row tile jt, image n), advanced incrementally
    struct Coord { int ct, jt, n; };
    auto decode = [&](int tile) {
        Coord c;
        c.ct = tile % p.col_tiles;
        c.jt = (tile / p.col_tiles) % row_tiles;
        c.n = tile / (p.col_tiles * row_tiles);
        return c;
    };
    auto advance = [&](Coord &c, int steps) {
        c.ct += steps;
        while (c.ct >= p.col_tiles) {
            c.ct -= p.col_tiles;
            if (++c.jt == row_tiles) { c.jt = 0; ++c.n; }
        }
    };
    const uint32_t patch_base = ptx::smem_u32(s_patch);
    const uint32_t raw_base = ptx::smem_u32(s_raw);

    if (warp < kProducerWarps) {
        // ------------------------------------------------------------ patch producers
        // One pixel per lane: 1360 patch pixels, thread handles pixels tid, tid + 256, ...  Lane L of a warp
        // then writes the 16-byte unit of pixel base + L, i.e. a warp store covers 512 contiguous bytes in the
        // minimum of 4 shared-memory wavefronts, and its three 4-byte loads at a 12-byte lane stride are
        // conflict-free too.  Everything that does not depend on the tile is hoisted: per pixel and tile the loop
        // is three shared loads, three adds (or table look-ups), one select per word and one 16-byte store, all
        // through explicit shared-space accesses (round 1's version went through generic pointers and predicated
        // blocks: three times the instructions, and these eight warps were the busiest role of the kernel).
        constexpr int kPx = kPatchRows * kPatchPx;                               // 1360
        constexpr int kPxRounds = (kPx + kProducerWarps * 32 - 1) / (kProducerWarps * 32);   // 6
        int px_row[kPxRounds], px_col[kPxRounds];
        uint32_t px_src[kPxRounds];
#pragma unroll
        for (int it = 0; it < kPxRounds; ++it) {
            const int i = threadIdx.x + kProducerWarps * 32 * it;
            const int r = i / kPatchPx, x = i - r * kPatchPx;
            px_row[it] = i < kPx ? r : 1 << 20;                     // pixels past the patch fail the row test below
            px_col[it] = x - 4;
            px_src[it] = i < kPx ? (IN_U8 ? r * kRawRowBytesU8 + (x - 4 + kRawLeadU8) * 3 : i * 12) : 0;
        }
        Coord tc = decode(tile_begin);
        for (int t = 0; t < my_tiles; ++t, advance(tc, 1)) {
            const int row0 = kTileRows * tc.jt - 1, w0 = tc.ct * 128;
            const int slot = t % kRawDepth, stage = t % kStages;
            ptx::mbar_wait(bar_rfull + 8 * slot, (t / kRawDepth) & 1);        // raw pixels have landed
            ptx::mbar_wait(bar_empty + 8 * stage, ((t / kStages) & 1) ^ 1);    // patch slot is free
            const uint32_t patch = patch_base + stage * kPatchBytes + threadIdx.x * 16;
            const uint32_t raw = raw_base + slot * kRawBytes;
#pragma unroll
            for (int it = 0; it < kPxRounds; ++it) {
                if (it < kPxRounds - 1 || px_row[it] < kPatchRows) {           // only the last round is partial
                    const bool ok = (unsigned)(row0 + px_row[it]) < (unsigned)p.H && (unsigned)(w0 + px_col[it]) < (unsigned)p.W;
                    uint32_t o0, o1, o2;
                    if constexpr (IN_U8) {
                        // the table holds the finished operand bits of float32(v / 255.)
                        const uint8_t *src = s_raw + slot * kRawBytes + px_src[it];
                        const uint32_t v0 = s_lut[src[0]], v1 = s_lut[src[1]], v2 = s_lut[src[2]];
                        const uint32_t c0 = p.swap_rb ? v2 : v0, c2 = p.swap_rb ? v0 : v2;
                        if constexpr (BF16) { o0 = c0 | (v1 << 16); o1 = c2; o2 = 0; }
                        else { o0 = c0; o1 = v1; o2 = c2; }
                    } else {
                        uint32_t f0, f1, f2;
                        const uint32_t a = raw + px_src[it];
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(f0) : "r"(a));
                        asm volatile("ld.shared.b32 %0, [%1+4];" : "=r"(f1) : "r"(a));
                        asm volatile("ld.shared.b32 %0, [%1+8];" : "=r"(f2) : "r"(a));
                        if constexpr (BF16) {
                            __nv_bfloat162 rg = __floats2bfloat162_rn(__uint_as_float(f0), __uint_as_float(f1));
                            __nv_bfloat162 b0 = __floats2bfloat162_rn(__uint_as_float(f2), 0.f);
                            o0 = *reinterpret_cast<uint32_t *>(&rg);
                            o1 = *reinterpret_cast<uint32_t *>(&b0);
                            o2 = 0;
                        } else {
                            // round to nearest TF32: add half an ulp; the MMA drops the low 13 bits itself
                            o0 = f0 + 0x1000u; o1 = f1 + 0x1000u; o2 = f2 + 0x1000u;
                        }
                    }
                    if (!ok) { o0 = 0; o1 = 0; o2 = 0; }                         // outside the image: zeros (SAME padding)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(patch + kProducerWarps * 32 * 16 * it), "r"(o0),
                                 "r"(o1), "r"(o2), "r"(0u) : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(bar_rempty + 8 * slot);      // raw slot may be refilled
                ptx::mbar_arrive(bar_full + 8 * stage);       // patch ready for the MMAs
            }
        }
    } else if (warp == kLoaderWarp) {
        // ------------------------------------------------------------ raw-pixel loader (one thread)
        if (ptx::elect_one()) {
            Coord tc = decode(tile_begin);
            for (int t = 0; t < my_tiles; ++t, advance(tc, 1)) {
                const int slot = t % kRawDepth;
                ptx::mbar_wait(bar_rempty + 8 * slot, ((t / kRawDepth) & 1) ^ 1);
                constexpr int lead = IN_U8 ? kRawLeadU8 : 4, span = IN_U8 ? kRawPxU8 : kPatchPx;
                constexpr int px_bytes = IN_U8 ? 3 : 12, row_bytes = IN_U8 ? kRawRowBytesU8 : kRawRowBytes;
                const int row0 = kTileRows * tc.jt - 1, w0 = tc.ct * 128;
                const int px_lo = w0 - lead > 0 ? w0 - lead : 0;
                const int px_hi = w0 + span - lead < p.W ? w0 + span - lead : p.W;
                const int r_lo = row0 < 0 ? -row0 : 0;
                const int r_hi = row0 + kPatchRows <= p.H ? kPatchRows : p.H - row0;
                const uint32_t bytes = (uint32_t)(px_hi - px_lo) * px_bytes;  // a multiple of 16 (pixels in fours / sixteens)
                ptx::mbar_expect_tx(bar_rfull + 8 * slot, bytes * (uint32_t)(r_hi - r_lo));
                const uint8_t *src = static_cast<const uint8_t *>(p.in) +
                                     (((size_t)tc.n * p.H + (row0 + r_lo)) * p.W + px_lo) * px_bytes;
                uint32_t dst = raw_base + slot * kRawBytes + r_lo * row_bytes + (px_lo - (w0 - lead)) * px_bytes;
                for (int r = r_lo; r < r_hi; ++r, src += (size_t)p.W * px_bytes, dst += row_bytes)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar_rfull + 8 * slot) : "memory");
            }
        }
    } else if (warp < kProducerWarps + kMmaWarps) {
        // ------------------------------------------------------------ MMA issuers
        if (ptx::elect_one()) {
            constexpr uint32_t fmt = BF16 ? 1u : 2u;
            // descriptor words: lo = start >> 4 | LBO >> 4 << 16, hi = SBO >> 4 | version; no swizzle.
            // A: SBO 128 (next 8 pixels); LBO 16 (next pixel = next tap) or one patch row (the same tap of the next row).
            // B: SBO 128; LBO = rows of the operand x 16 (48 rows: 768, 64 rows: 1024).
            constexpr uint32_t HI = (128u >> 4) | (1u << 14);
            const uint32_t w_lo0 = (ptx::smem_u32(s_w) >> 4) | ((768u >> 4) << 16);
            const uint32_t w2_lo0 = ((ptx::smem_u32(s_w) + kWOperandBytes) >> 4) | ((1024u >> 4) << 16);
            for (int t = warp - kProducerWarps; t < my_tiles; t += kMmaWarps) {
                const int stage = t % kStages, acc = t % kAccStages;
                ptx::mbar_wait(bar_tempty + 8 * acc, ((t / kAccStages) & 1) ^ 1);
                ptx::mbar_wait(bar_full + 8 * stage, (t / kStages) & 1);
                ptx::tc_fence_after();
                const uint32_t a_base = (patch_base + stage * kPatchBytes) >> 4;
                const uint32_t d = tmem_base + acc * kAccCols;
                // Taps (dj 0, dj 1) of patch row r feed output rows o in [max(0, r-2), min(7, r)] with di = r - o: the
                // stacked operand [W2|W1|W0] from row 16 (2 - (r - o_lo)), N = 16 (#rows), into columns 16 o_lo.
                // Rows 2, 5, 8 go first: together they touch every column exactly once and initialise it.
#pragma unroll
                for (int oi = 0; oi < kPatchRows; ++oi) {
                    const int r = oi < 3 ? 2 + 3 * oi : (oi == 3 ? 0 : oi == 4 ? 1 : oi == 5 ? 3 : oi == 6 ? 4 : oi == 7 ? 6 : oi == 8 ? 7 : 9);
                    const int o_lo = r - 2 > 0 ? r - 2 : 0;
                    const int o_hi = r < kTileRows - 1 ? r : kTileRows - 1;
                    const int n = 16 * (o_hi - o_lo + 1);
                    const int wrow = 16 * (2 - (r - o_lo));
                    // A rows: pixel m, taps dj = 0, 1  ->  patch pixels m + 3, m + 4
                    const uint32_t a_lo = a_base + ((r * kPatchRowBytes + 3 * 16) >> 4) + ((16u >> 4) << 16);
                    ptx::umma<BF16>(d + 16 * o_lo, ptx::umma_desc_join(a_lo, HI), ptx::umma_desc_join(w_lo0 + ((wrow * 16) >> 4), HI),
                                    ptx::umma_idesc(fmt, n), oi >= 3 ? 1u : 0u);
                }
                // Tap dj 2 of patch rows r and r + 1 (r even): K halves (row r, pixel m + 5) and (row r + 1, pixel m + 5),
                // output rows o in [max(0, r-2), min(7, r+1)].  Operand 1 has 64 rows, 16 (o - (r - 2)) + oc, with
                // W[r - o][2] in the first half and W[r + 1 - o][2] in the second (zero where di leaves 0..2).
#pragma unroll
                for (int r = 0; r < kPatchRows; r += 2) {
                    const int o_lo = r - 2 > 0 ? r - 2 : 0;
                    const int o_hi = r + 1 < kTileRows - 1 ? r + 1 : kTileRows - 1;
                    const int n = 16 * (o_hi - o_lo + 1);
                    const int wrow = 16 * (o_lo - (r - 2));
                    const uint32_t a_lo = a_base + ((r * kPatchRowBytes + 5 * 16) >> 4) + (((uint32_t)kPatchRowBytes >> 4) << 16);
                    ptx::umma<BF16>(d + 16 * o_lo, ptx::umma_desc_join(a_lo, HI), ptx::umma_desc_join(w2_lo0 + ((wrow * 16) >> 4), HI),
                                    ptx::umma_idesc(fmt, n), 1u);
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
        const int OW = p.W >> 1, OH = p.H >> 1;
        const bool odd = lane & 1;
        float sc8[8], sh8[8];                          // scale/shift of the 8 channels this lane stores
#pragma unroll
        for (int c = 0; c < 8; ++c) { sc8[c] = s_scale[(odd ? 8 : 0) + c]; sh8[c] = s_shift[(odd ? 8 : 0) + c]; }
        // The three epilogue variants are three separate tile loops, chosen once: inside one loop the compiler
        // if-converts the variants and every thread issues all of them (round 1's version: 1069 issue slots per
        // pooling row pair for ~100 useful ones -- after the converters the busiest role of the kernel).
        //   0: every scale > 0 -> pool the raw accumulators, then activate (the normal case)
        //   1: activate, then pool (negative scales)          2: as 1, and dump the pooled-over activations (debug)
        auto epilogue_tiles = [&](auto MODE) {
            constexpr int mode = decltype(MODE)::value;
            Coord tc = decode(tile_begin + group);
            for (int t = group; t < my_tiles; t += kEpilogueGroups, advance(tc, kEpilogueGroups)) {
                const int ct = tc.ct, n = tc.n, jt = tc.jt;
                const int w = ct * 128 + m;
                const bool valid = w < p.W;
                const int acc = t % kAccStages;
                ptx::mbar_wait(bar_tfull + 8 * acc, (t / kAccStages) & 1);
                ptx::tc_fence_after();
                if (ct * 128 + quarter * 32 < p.W) {      // warp-uniform: skip quarters right of the image
#pragma unroll 1
                    for (int rp = 0; rp < kTileRows / 2; ++rp) {      // pooling row pairs of the tile
                        uint32_t raw[32];                             // [0,16): output row 2rp, [16,32): row 2rp+1
                        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kAccCols + 32 * rp, raw);
                        ptx::tmem_ld_wait();
                        const int orow = kTileRows * jt + 2 * rp;     // conv output row of raw[0..15]
                        if constexpr (mode == 2) {                   // debug dump of the pooled-over activations (cold path)
                            for (int c = 0; c < 16; ++c) {
                                float a = fmaf(__uint_as_float(raw[c]), s_scale[c], s_shift[c]);
                                float b = fmaf(__uint_as_float(raw[16 + c]), s_scale[c], s_shift[c]);
                                if (p.leaky) { a = fmaxf(a, 0.1f * a); b = fmaxf(b, 0.1f * b); }
                                if constexpr (OUT_BF16) { a = __bfloat162float(__float2bfloat16_rn(a)); b = __bfloat162float(__float2bfloat16_rn(b)); }
                                else if (p.round_tf32) { a = ptx::round_tf32(a); b = ptx::round_tf32(b); }
                                if (valid) {
                                    p.prepool[(((size_t)n * p.H + orow) * p.W + w) * 16 + c] = a;
                                    p.prepool[(((size_t)n * p.H + orow + 1) * p.W + w) * 16 + c] = b;
                                }
                            }
                        }
                        float r8[8];
                        if constexpr (mode == 0) {
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
                            float v[16];
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                float a = fmaf(__uint_as_float(raw[c]), s_scale[c], s_shift[c]);
                                float b = fmaf(__uint_as_float(raw[16 + c]), s_scale[c], s_shift[c]);
                                if (p.leaky) { a = fmaxf(a, 0.1f * a); b = fmaxf(b, 0.1f * b); }
                                float mx = fmaxf(a, b);                      // vertical max
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
                        if (valid) {
                            const size_t pix = ((size_t)n * OH + (orow >> 1)) * OW + (w >> 1);
                            if constexpr (OUT_BF16) {
                                __nv_bfloat16 *o = ptx::out_at<__nv_bfloat16>(p.out, pix * 16 + (odd ? 8 : 0), 8, p.out_elems, 1);
                                __nv_bfloat162 t0 = __floats2bfloat162_rn(r8[0], r8[1]), t1 = __floats2bfloat162_rn(r8[2], r8[3]);
                                __nv_bfloat162 t2 = __floats2bfloat162_rn(r8[4], r8[5]), t3 = __floats2bfloat162_rn(r8[6], r8[7]);
                                uint4 pk;
                                pk.x = *reinterpret_cast<uint32_t *>(&t0); pk.y = *reinterpret_cast<uint32_t *>(&t1);
                                pk.z = *reinterpret_cast<uint32_t *>(&t2); pk.w = *reinterpret_cast<uint32_t *>(&t3);
                                *reinterpret_cast<uint4 *>(o) = pk;
                            } else {
                                float *o = ptx::out_at<float>(p.out, pix * 16 + (odd ? 8 : 0), 8, p.out_elems, 2);
                                *reinterpret_cast<float4 *>(o) = make_float4(r8[0], r8[1], r8[2], r8[3]);
                                *reinterpret_cast<float4 *>(o + 4) = make_float4(r8[4], r8[5], r8[6], r8[7]);
                            }
                        }
                    }
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
    if (warp == kProducerWarps) ptx::tmem_dealloc(tmem_base, kAccCols * kAccStages);
}

// Host side: the two stacked weight operands, each in the canonical no-swizzle K-major layout
//   elem(nrow, k) -> (k / KC) * 16 ROWS + (nrow / 8) * 128 + (nrow % 8) * 16 + (k % KC) * ES
// with KC = elements per 16 bytes and ROWS the operand's row count.
//   operand 0 (48 rows): rows 0-15 hold W[di=2], 16-31 W[di=1], 32-47 W[di=0] (row = 16 (2-di) + oc), so any patch
//     row's set of output rows is a contiguous run of operand rows; K halves = taps dj 0, dj 1.
//   operand 1 (64 rows), tap dj 2 of patch rows r (first K half) and r + 1 (second half): row = 16 j + oc stands for
//     output row o = r - 2 + j, j = 0..3; first half W[di = 2 - j][2] (j <= 2), second half W[di = 3 - j][2] (j >= 1).
// Channels c >= 3 are zero.  kernel: HWIO [3][3][3][16] fp32.  rnd(): value -> operand bit pattern.
template <typename Elem, typename Rnd>
void pack_weights(const float *kernel, Elem *out /* kWBytes */, Rnd rnd) {
    static_assert(sizeof(Elem) == 4 || sizeof(Elem) == 2, "fp32 (tf32) or bf16 operands");
    for (int i = 0; i < kWBytes / (int)sizeof(Elem); ++i) out[i] = rnd(0.f);
    auto put = [&](Elem *op, int rows, int nrow, int half, int di, int dj) {
        for (int c = 0; c < 3; ++c) {
            const float wv = kernel[((size_t)(di * 3 + dj) * 3 + c) * 16 + (nrow % 16)];
            const size_t byte = (size_t)half * rows * 16 + (nrow / 8) * 128 + (nrow % 8) * 16 + c * sizeof(Elem);
            op[byte / sizeof(Elem)] = rnd(wv);
        }
    };
    for (int di = 0; di < 3; ++di)
        for (int oc = 0; oc < 16; ++oc)
            for (int half = 0; half < 2; ++half) put(out, 48, 16 * (2 - di) + oc, half, di, half);
    Elem *op1 = out + kWOperandBytes / sizeof(Elem);
    for (int j = 0; j < 4; ++j)
        for (int oc = 0; oc < 16; ++oc) {
            if (j <= 2) put(op1, 64, 16 * j + oc, 0, 2 - j, 2);
            if (j >= 1) put(op1, 64, 16 * j + oc, 1, 3 - j, 2);
        }
}

}  // namespace conv0
