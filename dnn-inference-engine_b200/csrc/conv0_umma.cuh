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
// (N = 48, clipped to 32 / 16 at the tile's top and bottom) into columns starting at 16 max(0, r-2):
// 20 MMAs per 1024 output pixels.  The epilogue thread of lane m holds all 8 rows of pixel column
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
constexpr int kWOperandBytes = 2 * 48 * 16;        // one stacked operand: 2 K halves x 48 rows x 16 B
constexpr int kWBytes = 2 * kWOperandBytes;        // tap pairs (dj 0,1) and (dj 2, -)
constexpr int kRawDepth = 6;                       // raw-pixel staging ring: tiles whose global loads are in flight
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
    int col_tiles;       // column tiles per row tile: ceil(W / 128), or W / 128 when the narrow remainder is packed
    int num_tiles;       // B * tiles_per_image
    // Packed narrow tiles (pack >= 2): W = 128 col_tiles + narrow.  Instead of a column tile that carries only
    // `narrow` valid pixels through 128 MMA rows, the narrow right-hand parts of `pack` consecutive row tiles
    // share one tile: segment s = MMA rows [seg_pitch s, seg_pitch s + narrow), seg_pitch = narrow + 2 (each
    // segment has its own one-pixel halo in the patch).  An image is then (H/8) col_tiles ordinary tiles
    // followed by ceil((H/8) / pack) packed ones.  416 columns: 174 tiles per image instead of 208.
    int pack, seg_pitch, narrow, tiles_per_image;
    int leaky, round_tf32;
    int pool_first;      // every scale[c] > 0: scale/shift, LeakyReLU and rounding are then monotone, so the
                         // epilogue pools the raw accumulators and activates one value instead of four
    float *prepool;      // debug: also dump the pre-pool activations [B,H,W,16] fp32
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
    // tile -> (column tile ct, row tile jt, image n), advanced incrementally.  With packing, ct == col_tiles
    // marks a packed tile whose first row tile is jt.
    struct Coord { int ct, jt, n; };
    auto decode = [&](int tile) {
        Coord c;
        c.n = tile / p.tiles_per_image;
        const int t = tile - c.n * p.tiles_per_image;
        if (t < row_tiles * p.col_tiles) {
            c.ct = t % p.col_tiles;
            c.jt = t / p.col_tiles;
        } else {
            c.ct = p.col_tiles;
            c.jt = (t - row_tiles * p.col_tiles) * p.pack;
        }
        return c;
    };
    auto advance = [&](Coord &c, int steps) {
        for (int i = 0; i < steps; ++i) {
            if (c.ct < p.col_tiles) {
                if (++c.ct == p.col_tiles) {
                    c.ct = 0;
                    if (++c.jt == row_tiles) {
                        c.jt = 0;
                        if (p.pack) c.ct = p.col_tiles;       // the image's packed tiles follow
                        else ++c.n;
                    }
                }
            } else {
                c.jt += p.pack;
                if (c.jt >= row_tiles) { c.ct = 0; c.jt = 0; ++c.n; }
            }
        }
    };
    const int w0_packed = p.col_tiles * 128;

    if (warp < kProducerWarps) {
        // ------------------------------------------------------------ patch producers
        // One pixel per lane: 1360 patch pixels, thread handles pixels tid, tid + 256, ...  Lane L of a warp
        // then writes the 16-byte unit of pixel base + L, i.e. a warp store covers 512 contiguous bytes in the
        // minimum of 4 shared-memory wavefronts (writing a 4-pixel group per thread, lanes 64 bytes apart,
        // cost 16: measured, more than half of all the kernel's shared-memory wavefronts were store
        // conflicts), and its three 4-byte loads at a 12-byte lane stride are conflict-free too.
        // The pixel -> (patch row, column) split does not depend on the tile: hoist it.
        constexpr int kPx = kPatchRows * kPatchPx;                               // 1360
        constexpr int kPxRounds = (kPx + kProducerWarps * 32 - 1) / (kProducerWarps * 32);   // 6
        int px_row[kPxRounds], px_col[kPxRounds], px_src[kPxRounds];
#pragma unroll
        for (int it = 0; it < kPxRounds; ++it) {
            const int i = threadIdx.x + kProducerWarps * 32 * it;
            const int r = i / kPatchPx, x = i - r * kPatchPx;
            px_row[it] = i < kPx ? r : 1 << 20;                     // pixels past the patch fail the row test below
            px_col[it] = x - 4;
            px_src[it] = IN_U8 ? r * kRawRowBytesU8 + (x - 4 + kRawLeadU8) * 3 : i * 12;
        }
        Coord tc = decode(tile_begin);
        for (int t = 0; t < my_tiles; ++t, advance(tc, 1)) {
            const int row0 = kTileRows * tc.jt - 1, w0 = tc.ct * 128;
            const int slot = t % kRawDepth, stage = t % kStages;
            ptx::mbar_wait(bar_rfull + 8 * slot, (t / kRawDepth) & 1);        // raw pixels have landed
            ptx::mbar_wait(bar_empty + 8 * stage, ((t / kStages) & 1) ^ 1);    // patch slot is free
            uint4 *patch = reinterpret_cast<uint4 *>(s_patch + stage * kPatchBytes);
            const uint8_t *raw = s_raw + slot * kRawBytes;
            const bool packed = tc.ct == p.col_tiles;
#pragma unroll
            for (int it = 0; it < kPxRounds; ++it) {
                if (px_row[it] < kPatchRows) {
                    bool ok;
                    int src_off;
                    if (!packed) {
                        ok = (unsigned)(row0 + px_row[it]) < (unsigned)p.H && (unsigned)(w0 + px_col[it]) < (unsigned)p.W;
                        src_off = px_src[it];
                    } else {
                        // patch pixel x = px_col + 4 belongs to segment (x - 3) / seg_pitch, its column cx there is
                        // image column w0 - 1 + cx of row tile jt + segment; raw rows are stored [segment][row]
                        constexpr int lead = IN_U8 ? kRawLeadU8 : 4, px_bytes = IN_U8 ? 3 : 12;
                        const int xs = px_col[it] + 1;
                        const int sg = xs >= 0 ? xs / p.seg_pitch : p.pack;
                        const int cx = xs - sg * p.seg_pitch;
                        const int irow = kTileRows * (tc.jt + sg) - 1 + px_row[it];
                        ok = sg < p.pack && tc.jt + sg < row_tiles && (unsigned)irow < (unsigned)p.H &&
                             (unsigned)(w0_packed - 1 + cx) < (unsigned)p.W;
                        src_off = ((sg * kPatchRows + px_row[it]) * (p.narrow + 2 * lead) + cx - 1 + lead) * px_bytes;
                    }
                    uint4 o = make_uint4(0u, 0u, 0u, 0u);                        // outside the image: zeros (SAME padding)
                    if (ok) {
                        if constexpr (IN_U8) {
                            // the table holds the finished operand bits of float32(v / 255.)
                            const uint8_t *src = raw + src_off;
                            const uint32_t v0 = s_lut[src[0]], v1 = s_lut[src[1]], v2 = s_lut[src[2]];
                            const uint32_t c0 = p.swap_rb ? v2 : v0, c2 = p.swap_rb ? v0 : v2;
                            if constexpr (BF16) { o.x = c0 | (v1 << 16); o.y = c2; }
                            else { o.x = c0; o.y = v1; o.z = c2; }
                        } else {
                            const uint32_t *src = reinterpret_cast<const uint32_t *>(raw + src_off);
                            const uint32_t f0 = src[0], f1 = src[1], f2 = src[2];
                            if constexpr (BF16) {
                                __nv_bfloat162 rg = __floats2bfloat162_rn(__uint_as_float(f0), __uint_as_float(f1));
                                __nv_bfloat162 b0 = __floats2bfloat162_rn(__uint_as_float(f2), 0.f);
                                o.x = *reinterpret_cast<uint32_t *>(&rg);
                                o.y = *reinterpret_cast<uint32_t *>(&b0);
                            } else {
                                // round to nearest TF32: add half an ulp; the MMA drops the low 13 bits itself
                                o.x = f0 + 0x1000u; o.y = f1 + 0x1000u; o.z = f2 + 0x1000u;
                            }
                        }
                    }
                    patch[threadIdx.x + kProducerWarps * 32 * it] = o;
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
            const uint32_t raw_base = ptx::smem_u32(s_raw);
            Coord tc = decode(tile_begin);
            for (int t = 0; t < my_tiles; ++t, advance(tc, 1)) {
                const int slot = t % kRawDepth;
                ptx::mbar_wait(bar_rempty + 8 * slot, ((t / kRawDepth) & 1) ^ 1);
                constexpr int lead = IN_U8 ? kRawLeadU8 : 4, span = IN_U8 ? kRawPxU8 : kPatchPx;
                constexpr int px_bytes = IN_U8 ? 3 : 12, row_bytes = IN_U8 ? kRawRowBytesU8 : kRawRowBytes;
                if (tc.ct == p.col_tiles) {
                    // packed tile: per segment (row tile jt + sg) the rows of columns [w0 - lead, W), stored [segment][row]
                    const int nseg = min(p.pack, row_tiles - tc.jt);
                    const uint32_t bytes = (uint32_t)(p.W - (w0_packed - lead)) * px_bytes;     // a multiple of 16
                    const int seg_row_bytes = (p.narrow + 2 * lead) * px_bytes;
                    uint32_t rows = 0;
                    for (int sg = 0; sg < nseg; ++sg) {
                        const int row0 = kTileRows * (tc.jt + sg) - 1;
                        rows += (uint32_t)((row0 + kPatchRows <= p.H ? kPatchRows : p.H - row0) - (row0 < 0 ? -row0 : 0));
                    }
                    ptx::mbar_expect_tx(bar_rfull + 8 * slot, bytes * rows);
                    for (int sg = 0; sg < nseg; ++sg) {
                        const int row0 = kTileRows * (tc.jt + sg) - 1;
                        const int r_lo = row0 < 0 ? -row0 : 0;
                        const int r_hi = row0 + kPatchRows <= p.H ? kPatchRows : p.H - row0;
                        const uint8_t *src = static_cast<const uint8_t *>(p.in) +
                                             (((size_t)tc.n * p.H + (row0 + r_lo)) * p.W + (w0_packed - lead)) * px_bytes;
                        uint32_t dst = raw_base + slot * kRawBytes + (sg * kPatchRows + r_lo) * seg_row_bytes;
                        for (int r = r_lo; r < r_hi; ++r, src += (size_t)p.W * px_bytes, dst += seg_row_bytes)
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(dst), "l"(src), "r"(bytes), "r"(bar_rfull + 8 * slot) : "memory");
                    }
                    continue;
                }
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
            // A: LBO 16 (next pixel = next tap), SBO 128 (next 8 pixels).  B: LBO 768 (48 rows), SBO 128.
            constexpr uint32_t HI = (128u >> 4) | (1u << 14);
            const uint32_t w_lo0 = (ptx::smem_u32(s_w) >> 4) | ((768u >> 4) << 16);
            for (int t = warp - kProducerWarps; t < my_tiles; t += kMmaWarps) {
                const int stage = t % kStages, acc = t % kAccStages;
                ptx::mbar_wait(bar_tempty + 8 * acc, ((t / kAccStages) & 1) ^ 1);
                ptx::mbar_wait(bar_full + 8 * stage, (t / kStages) & 1);
                ptx::tc_fence_after();
                const uint32_t a_lo0 = (ptx::smem_u32(s_patch + stage * kPatchBytes) >> 4) | ((16u >> 4) << 16);
                const uint32_t d = tmem_base + acc * kAccCols;
                // Patch row r feeds output rows o in [max(0, r-2), min(7, r)] with di = r - o: the stacked
                // operand [W2|W1|W0] from row 16 (2 - (r - o_lo)), N = 16 (#rows), into columns 16 o_lo.
                // Rows 2, 5, 8 go first: together they touch every column once and initialise it.
#pragma unroll
                for (int oi = 0; oi < kPatchRows; ++oi) {
                    const int r = oi < 3 ? 2 + 3 * oi : (oi == 3 ? 0 : oi == 4 ? 1 : oi == 5 ? 3 : oi == 6 ? 4 : oi == 7 ? 6 : oi == 8 ? 7 : 9);
                    const int o_lo = r - 2 > 0 ? r - 2 : 0;
                    const int o_hi = r < kTileRows - 1 ? r : kTileRows - 1;
                    const int n = 16 * (o_hi - o_lo + 1);
                    const int wrow = 16 * (2 - (r - o_lo));
                    const uint32_t idesc = ptx::umma_idesc(fmt, n);
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr) {
                        // A rows: pixel m, taps dj = 2*pr, 2*pr+1  ->  patch pixel m + dj + 3
                        const uint32_t a_off = (r * kPatchRowBytes + (3 + 2 * pr) * 16) >> 4;
                        const uint32_t b_off = (pr * kWOperandBytes + wrow * 16) >> 4;
                        ptx::umma<BF16>(d + 16 * o_lo, ptx::umma_desc_join(a_lo0 + a_off, HI), ptx::umma_desc_join(w_lo0 + b_off, HI),
                                        idesc, (oi >= 3 || pr != 0) ? 1u : 0u);
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
        const int OW = p.W >> 1, OH = p.H >> 1;
        const bool odd = lane & 1;
        float sc8[8], sh8[8];                          // scale/shift of the 8 channels this lane stores
#pragma unroll
        for (int c = 0; c < 8; ++c) { sc8[c] = s_scale[(odd ? 8 : 0) + c]; sh8[c] = s_shift[(odd ? 8 : 0) + c]; }
        Coord tc = decode(tile_begin + group);
        for (int t = group; t < my_tiles; t += kEpilogueGroups, advance(tc, kEpilogueGroups)) {
            const int ct = tc.ct, n = tc.n;
            int jt = tc.jt, w = ct * 128 + m;
            bool valid = w < p.W;
            const bool packed = ct == p.col_tiles;
            if (packed) {                              // lane m = segment m / seg_pitch, column m % seg_pitch of row tile jt + segment
                const int sg = m / p.seg_pitch, cx = m - sg * p.seg_pitch;
                jt += sg;
                w = w0_packed + cx;
                valid = sg < p.pack && cx < p.narrow && jt < row_tiles;
            }
            const int acc = t % kAccStages;
            ptx::mbar_wait(bar_tfull + 8 * acc, (t / kAccStages) & 1);
            ptx::tc_fence_after();
            if (packed || ct * 128 + quarter * 32 < p.W) {      // warp-uniform: skip quarters right of the image
#pragma unroll 1
                for (int rp = 0; rp < kTileRows / 2; ++rp) {      // pooling row pairs of the tile
                    uint32_t raw[32];                             // [0,16): output row 2rp, [16,32): row 2rp+1
                    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kAccCols + 32 * rp, raw);
                    ptx::tmem_ld_wait();
                    const int orow = kTileRows * jt + 2 * rp;     // conv output row of raw[0..15]
                    if (p.prepool != nullptr) {                   // debug dump of the pooled-over activations (cold path)
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
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(bar_tempty + 8 * acc);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kProducerWarps) ptx::tmem_dealloc(tmem_base, kAccCols * kAccStages);
}

// Host side: the two stacked weight operands (tap pair pr = 0: dj 0,1; pr = 1: dj 2 and a zero tap),
// each in the canonical no-swizzle K-major layout
//   elem(nrow, k) -> (k / KC) * 768 + (nrow / 8) * 128 + (nrow % 8) * 16 + (k % KC) * ES
// with KC = elements per 16 bytes.  Rows 0-15 hold W[di=2], 16-31 W[di=1], 32-47 W[di=0] (row = 16 (2-di) + oc),
// so any patch row's set of output rows is a contiguous run of operand rows.  K index = (dj - 2 pr) KC + c;
// channels c >= 3 and tap dj = 3 are zero.  kernel: HWIO [3][3][3][16] fp32.
// rnd(): value -> operand bit pattern (tf32-rounded fp32 or bf16).
template <typename Elem, typename Rnd>
void pack_weights(const float *kernel, Elem *out /* kWBytes */, Rnd rnd) {
    static_assert(sizeof(Elem) == 4 || sizeof(Elem) == 2, "fp32 (tf32) or bf16 operands");
    for (int i = 0; i < kWBytes / (int)sizeof(Elem); ++i) out[i] = rnd(0.f);
    for (int pr = 0; pr < 2; ++pr) {
        Elem *op = out + (size_t)pr * kWOperandBytes / sizeof(Elem);
        for (int di = 0; di < 3; ++di)
            for (int oc = 0; oc < 16; ++oc) {
                const int nrow = 16 * (2 - di) + oc;
                for (int half = 0; half < 2; ++half) {
                    const int dj = 2 * pr + half;
                    if (dj > 2) continue;
                    for (int c = 0; c < 3; ++c) {
                        const float wv = kernel[((size_t)(di * 3 + dj) * 3 + c) * 16 + oc];
                        const size_t byte = (size_t)half * 768 + (nrow / 8) * 128 + (nrow % 8) * 16 + c * sizeof(Elem);
                        op[byte / sizeof(Elem)] = rnd(wv);
                    }
                }
            }
    }
}

}  // namespace conv0
