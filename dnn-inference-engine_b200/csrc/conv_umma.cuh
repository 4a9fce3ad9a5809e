// conv_umma.cuh -- the hot kernel: NHWC stride-1 convolution as an implicit GEMM on the
// sm_100a tensor cores (tcgen05.mma, fp32 accumulators in TMEM), operands staged by TMA,
// with bias + folded BatchNorm + LeakyReLU(0.1) (+ the 2x2/2 max-pool) fused in the epilogue.
// Replaces, for one conv block of proj3/yolov2tiny.py:30-77, the reference's
// np.pad -> im2col -> sgemm -> bias_add -> batch_norm -> leaky_relu -> np.pad -> max_pool2d
// chain (proj3/dnn_openblas.py:174-284, proj3/dnn_openblas.c:9-254).
//
// GEMM view: D[m][n] = sum_k A[m][k] * Wt[n][k], m = output pixel, n = output channel,
// k = (filter tap, input channel).  One K-block = one tap x BKC channels, where BKC channels
// are SWZ bytes (one swizzle span) of an NHWC pixel.  No im2col matrix ever exists: the A tile
// of a K-block is fetched straight from the activation tensor by one TMA request,
//   * tiled mode : a [nb, th, tw, BKC] box whose (h, w) origin is shifted by the tap; rows that
//                  fall outside the image are zero-filled by TMA = SAME padding.  The 128 rows
//                  of a tile form whole 2x2 pooling windows, so the pool is fused.
//   * im2col mode: 128 consecutive output pixels (running over rows and images) through a
//                  cuTensorMapEncodeIm2col descriptor; used for the 13x13 layers, whose M
//                  would otherwise not tile without waste.
// Weights are prepacked once as [tap][channel chunk][OCp][BKC] (K-major), so the B tile of a
// K block is one contiguous TMA box.
//
// CTA = 192 threads, persistent over tiles (static round-robin):
//   warp 0   : TMA producer (one elected lane)          full/empty mbarrier ring, STAGES deep
//   warp 1   : TMEM allocator + MMA issuer (one lane)   tcgen05.mma cta_group::1, M=128, N=BN
//   warps 2-5: epilogue, one TMEM lane quarter each     2 accumulator stages (tmem_full/empty)
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda.h>
#include <cuda_bf16.h>

#include "epilogue_pool1.cuh"
#include "ptx.cuh"

namespace umma_conv {

constexpr int kThreads = 192;
constexpr int kBM = 128;

struct Params {
    int H, W, B;          // conv output = input spatial size (stride 1, SAME)
    int OC;               // real output channels (columns beyond are dropped)
    int KH, KW, PT, PL;   // filter and top/left padding
    int c_chunks;         // K-blocks per tap = ICp / BKC
    int num_m_tiles, num_n_tiles;
    // tiled mode: tile = nb images x th rows x tw cols (all powers of two, product 128)
    int tw, th, nb, log2_tw, log2_th, tiles_w, tiles_h;
    int M_total;          // im2col mode: B*H*W
    int m_base;           // im2col mode: first output pixel of this launch (a block may be cut into a main launch
                          // of whole waves and a finer-tiled tail launch over the remaining rows)
    int m_stride;         // im2col mode: output pixels a 128-row tile advances: 128; with a pool fused through the staging
                          // area (POOL1, epilogue_pool1.cuh) 128 - (W + 1) for the 2x2/1 SAME pool (the last W + 1 rows of
                          // a tile are halo) or a whole number of image-row pairs (2kW) for the 2x2/2 pool
    int pool_stride;      // POOL1: 1 = 2x2/1 SAME pool, 2 = 2x2/2 pool
    const float *scale;   // [num_n_tiles*BN]  gamma / sqrt(var + eps)      (1 when no BN)
    const float *shift;   // [num_n_tiles*BN]  (bias - mean) * scale         (bias when no BN)
    void *out;            // [B, OH, OW, OC]
    int OH, OW;           // stored spatial size (H/2, W/2 when the pool is fused)
    int leaky;
    int round_tf32;       // fp32 storage only: round stored values to TF32 (they feed a kind::tf32 block)
    float *prepool;       // debug (fused pool only): also dump the pre-pool activations [B,H,W,OC] fp32
    int tma_store;        // STAGE_OUT: write the staged rows with bulk async copies (0: per-lane stores; DNNB_NO_TMA_STORE)
    int pool_first;       // POOL: every scale[c] > 0 -> the launcher picks the POOL_FIRST instantiation
    // split-K (small batches: too few output tiles to fill the SMs, weights dominate the traffic):
    // `splits` CTAs share one output tile.  Each stores the partial sums of its K slice to its own
    // slab of `workspace`, bumps `counters[tile]` and waits until all `splits` slabs are there; then
    // every CTA sums a 1/splits share of the tile's rows over the slabs IN SLAB ORDER and runs the
    // epilogue on it: deterministic, no atomics on data.  All CTAs of the launch must be co-resident
    // (grid = work items <= SMs).  A tile has an arrival and a departure counter; the last CTA to leave
    // the wait zeroes both, so every launch starts from zero without host bookkeeping (the launch sequence
    // can be replayed from a CUDA graph).
    int splits;           // 1 = off
    float *workspace;     // [splits][num_m_tiles * MT * 128][num_n_tiles * BN] fp32
    int *counters;        // [num_m_tiles * num_n_tiles][2]: arrived, departed; zero between launches
    // Cluster split-K (cluster_splitk = 1, splits in {2, 4, 8, 16}): the `splits` CTAs of an output tile are ONE thread-block
    // cluster (launch attribute; CTA rank = K slice, so co-scheduling is the hardware's guarantee, not an assumption
    // about the grid).  Each CTA parks its partial tile in its own shared memory (the ring is idle by then), the
    // cluster synchronises, and every CTA sums a BN / splits wide column strip of the tile over its peers IN RANK ORDER
    // through distributed shared memory, then finishes and stores that strip.  No workspace in global memory, no
    // counters, no spinning -- and rows stay together in one CTA, which is what a fused pool needs.
    int cluster_splitk;
    // debug: per-CTA phase stamps (%globaltimer, ns) [gridDim.x][kTimelineSlots]: 0 entry, 1 prologue done, 2 predecessor
    // complete (griddepcontrol.wait returned), 3 first operands landed, 4 last MMA issued, 5 accumulator complete, 6 partial
    // tile parked / workspace written and all siblings seen, 7 result stored; split-K only: 8 own partial tile written to
    // the workspace (workspace form), 9 siblings' partial sums read and reduced (cluster form) / landed (workspace form).
    // NULL in production.
    unsigned long long *timeline;
    size_t out_elems;     // extent of `out` in elements (checked by -DDNNB_BOUNDS builds)
};

constexpr int kTimelineSlots = 10;
__device__ __forceinline__ void stamp(const Params &p, int slot) {
    if (p.timeline != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.timeline[(size_t)blockIdx.x * kTimelineSlots + slot] = t;
    }
}

// MT = M tiles (of 128 rows) per CTA tile that share every B tile: MT = 2 turns the CTA tile
// into 256 x BN, a third less L2->SMEM traffic per flop, at the price of a single accumulator
// stage when 2 * BN = 512 (the epilogue then no longer overlaps the next tile's MMAs; with the
// long K loops of the 13x13 layers that costs ~2%).
// STAGE_OUT reserves 4 x 32 rows x BN floats so that a tile whose rows are narrower than a 16-byte
// multiple (block 8: 125 channels = 500-byte rows) is staged in shared memory and written as one
// contiguous, coalesced run instead of 4-byte scattered stores.
template <int BN, int SWZ, int KBS, int MT = 1, bool STAGE_OUT = false, bool POOL1 = false>
struct Cfg {
    static constexpr int A_TILE = kBM * SWZ;
    static constexpr int B_TILE = BN * SWZ;
    static constexpr int STAGE_BYTES = KBS * (MT * A_TILE + B_TILE);
    static constexpr int OUT_STAGE_BYTES = STAGE_OUT ? kBM * BN * 4 : POOL1 ? pool1::kStageBytes : 0;
    static constexpr int TAIL_BYTES = 1024 /*barriers+tmem slot*/ + 2 * 1024 * 4 /*scale, shift*/ + OUT_STAGE_BYTES;
    static constexpr int SMEM_BUDGET = 227 * 1024 - 1024 /*alignment slack*/;
    static constexpr int STAGES_RAW = (SMEM_BUDGET - TAIL_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + TAIL_BYTES + 1024;
    static constexpr int ACC_COLS = MT * BN;
    static constexpr int ACC_STAGES = 2 * ACC_COLS <= 512 ? 2 : 1;
    static constexpr int TMEM_COLS = ACC_STAGES * ACC_COLS < 32 ? 32 : ACC_STAGES * ACC_COLS;
    static_assert(STAGES >= 2, "pipeline needs at least two stages");
    static_assert(ACC_COLS <= 512, "accumulators must fit TMEM");
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM allocations are powers of two");
};

template <bool BF16, bool OUT_BF16, int BN, int SWZ, int KBS, bool IM2COL, bool POOL, int MT = 1, bool STAGE_OUT = false,
          bool POOL1 = false, bool POOL_FIRST = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const Params p) {
    using C = Cfg<BN, SWZ, KBS, MT, STAGE_OUT, POOL1>;
    static_assert(!STAGE_OUT || (IM2COL && !POOL && !OUT_BF16 && MT == 1), "staged output: plain fp32 im2col tiles only");
    static_assert(!POOL1 || (IM2COL && !POOL && !STAGE_OUT), "the fused 2x2/1 pool needs im2col tiles");
    static_assert(!POOL_FIRST || POOL, "pool-first is a variant of the fused 2x2/2 pool");
    constexpr int STAGES = C::STAGES;
    constexpr int ACC = C::ACC_STAGES;
    constexpr int ES = BF16 ? 2 : 4;
    constexpr int BKC = SWZ / ES;            // channels per K-block
    constexpr int MMA_PER_KB = SWZ / 32;     // each MMA consumes 32 bytes of K (8 tf32 / 16 bf16)
    static_assert(!(POOL && IM2COL), "the fused pool needs spatial tiles");
    constexpr bool kClusterCapable = IM2COL && !POOL && !STAGE_OUT && MT == 1;     // variants that carry the cluster split-K path

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *tail = smem + STAGES * C::STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail);           // full[S], empty[S], tfull[2], tempty[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tail + 512);
    float *s_scale = reinterpret_cast<float *>(tail + 1024);
    float *s_shift = s_scale + 1024;
    float *s_out = s_shift + 1024;       // STAGE_OUT: [4 warps][32 rows][OC]; POOL1: [128][33] staging of one column chunk

    const uint32_t smem_base = ptx::smem_u32(smem);
    const uint32_t bar_full = ptx::smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * STAGES;
    const uint32_t bar_tempty = bar_tfull + 16;
    const uint32_t bar_red = bar_tempty + 16;        // split-K: bulk copies of the sibling slabs into the idle smem ring

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
    const int num_tiles = tiles_mn * p.splits;       // work items: (output tile, K slice)
    const int taps = p.KH * p.KW;
    const int num_kstages = taps * p.c_chunks / KBS;
    // work item -> (output tile, K slice).  Global-workspace split-K orders the items slice-major; cluster split-K
    // makes the slices of a tile consecutive CTAs (= the ranks of one cluster).
    auto item_split = [&](int tile) { return p.cluster_splitk ? tile % p.splits : tile / tiles_mn; };
    auto item_mn = [&](int tile) { return p.cluster_splitk ? tile / p.splits : tile % tiles_mn; };
    // K slice of work item `tile`: stages [ks_lo, ks_hi)
    auto k_range = [&](int tile, int &ks_lo, int &ks_hi) {
        const int split = item_split(tile);
        ks_lo = (int)((long)num_kstages * split / p.splits);
        ks_hi = (int)((long)num_kstages * (split + 1) / p.splits);
    };

    ptx::griddep_launch();
    if (threadIdx.x == 0) stamp(p, 0);
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(bar_full + 8 * s, 1);
            ptx::mbar_init(bar_empty + 8 * s, 1);
        }
        for (int s = 0; s < ACC; ++s) {
            ptx::mbar_init(bar_tfull + 8 * s, 1);
            ptx::mbar_init(bar_tempty + 8 * s, 128);
        }
        ptx::mbar_init(bar_red, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), C::TMEM_COLS);
    {   // epilogue constants: a CTA with a single work item only needs the columns of its own N tile
        const bool one_tile = num_tiles <= (int)gridDim.x;
        const int c_lo = one_tile ? (item_mn((int)blockIdx.x) % p.num_n_tiles) * BN : 0;
        const int c_hi = one_tile ? c_lo + BN : p.num_n_tiles * BN;
        for (int i = c_lo + threadIdx.x; i < c_hi; i += kThreads) {
            s_scale[i] = p.scale[i];
            s_shift[i] = p.shift[i];
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 32) stamp(p, 1);
    // Programmatic dependent launch: up to here nothing depends on the preceding kernel.  The producer
    // lane additionally starts the weight tiles of its first stages before it waits.
    if (warp != 0) ptx::griddep_wait();
    if (threadIdx.x == 32) stamp(p, 2);

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            // weights do not depend on the preceding kernel: arm the first ring pass and start its B tiles
            // before waiting for that kernel (the first pass over the ring finds every slot empty)
            int early = 0;
            if ((int)blockIdx.x < num_tiles) {
                int ks_lo, ks_hi;
                k_range(blockIdx.x, ks_lo, ks_hi);
                early = ks_hi - ks_lo < STAGES ? ks_hi - ks_lo : STAGES;
                const int n_tile = item_mn((int)blockIdx.x) % p.num_n_tiles;
                for (int e = 0; e < early; ++e) {
                    const uint32_t full = bar_full + 8 * e;
                    ptx::mbar_expect_tx(full, C::STAGE_BYTES);
                    const uint32_t sb = smem_base + e * C::STAGE_BYTES + KBS * MT * C::A_TILE;
#pragma unroll
                    for (int j = 0; j < KBS; ++j) {
                        const int kb = (ks_lo + e) * KBS + j;
                        const int tap = kb / p.c_chunks;
                        ptx::tma_load_4d(sb + j * C::B_TILE, &map_b, full, 0, n_tile * BN, kb - tap * p.c_chunks, tap);
                    }
                }
            }
            ptx::griddep_wait();
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n_tile = item_mn(tile) % p.num_n_tiles;
                const int m_tile = item_mn(tile) / p.num_n_tiles;
                int ks_lo, ks_hi;
                k_range(tile, ks_lo, ks_hi);
                int w0[MT], h0[MT], n0[MT];
                if constexpr (IM2COL) {
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        const int m0 = p.m_base + (m_tile * MT + i) * p.m_stride;
                        n0[i] = m0 / (p.H * p.W);
                        const int rem = m0 - n0[i] * (p.H * p.W);
                        h0[i] = rem / p.W;
                        w0[i] = rem - h0[i] * p.W;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < MT; ++i) {     // the MT sub-tiles are consecutive groups of nb images
                        w0[i] = (m_tile % p.tiles_w) * p.tw;
                        h0[i] = ((m_tile / p.tiles_w) % p.tiles_h) * p.th;
                        n0[i] = ((m_tile / (p.tiles_w * p.tiles_h)) * MT + i) * p.nb;
                    }
                }
                // (tap, channel chunk) of the next K block, advanced incrementally: one elected thread issues every copy of
                // the CTA, and two integer divisions per K block were a third of its instructions (profiles/r02_SUMMARY.md)
                int tap = (ks_lo * KBS) / p.c_chunks, cc = ks_lo * KBS - tap * p.c_chunks;
                int di = tap / p.KW, dj = tap - di * p.KW;
                for (int ks = ks_lo; ks < ks_hi; ++ks) {
                    const bool armed = early > 0;        // slot armed and its B tiles in flight since before the wait
                    if (armed) --early;
                    else ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t full = bar_full + 8 * stage;
                    if (!armed) ptx::mbar_expect_tx(full, C::STAGE_BYTES);
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + KBS * MT * C::A_TILE;
#pragma unroll
                    for (int j = 0; j < KBS; ++j) {
#pragma unroll
                        for (int i = 0; i < MT; ++i) {
                            const uint32_t dst = sa + (j * MT + i) * C::A_TILE;
                            if constexpr (IM2COL)
                                ptx::tma_load_im2col_4d(dst, &map_a, full, cc * BKC, w0[i] - p.PL, h0[i] - p.PT, n0[i],
                                                        (uint16_t)dj, (uint16_t)di);
                            else
                                ptx::tma_load_4d(dst, &map_a, full, cc * BKC, w0[i] + dj - p.PL, h0[i] + di - p.PT, n0[i]);
                        }
                        if (!armed) ptx::tma_load_4d(sb + j * C::B_TILE, &map_b, full, 0, n_tile * BN, cc, tap);
                        if (++cc == p.c_chunks) {
                            cc = 0;
                            ++tap;
                            if (++dj == p.KW) { dj = 0; ++di; }
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        if constexpr (kClusterCapable) {       // the cluster barriers of the split-K reduction count every thread of every CTA
            if (p.cluster_splitk && p.splits > 1 && (int)blockIdx.x < num_tiles) { __syncwarp(); ptx::cluster_sync_all(); ptx::cluster_sync_all(); }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::umma_idesc(BF16 ? 1u : 2u, BN);
            constexpr uint32_t HI = ptx::umma_desc_hi(SWZ);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
                const int acc = local % ACC;
                const uint32_t acc_phase = (local / ACC) & 1;
                ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
                int ks_lo, ks_hi;
                k_range(tile, ks_lo, ks_hi);
                for (int ks = ks_lo; ks < ks_hi; ++ks) {
                    ptx::mbar_wait(bar_full + 8 * stage, phase);
                    ptx::tc_fence_after();
                    if (ks == ks_lo && local == 0) stamp(p, 3);
                    const uint32_t a_lo = ptx::umma_desc_lo(smem_base + stage * C::STAGE_BYTES);
                    const uint32_t b_lo = a_lo + ((KBS * MT * C::A_TILE) >> 4);
#pragma unroll
                    for (int j = 0; j < KBS; ++j) {
#pragma unroll
                        for (int i = 0; i < MT; ++i) {
#pragma unroll
                            for (int k = 0; k < MMA_PER_KB; ++k)
                                ptx::umma<BF16>(d_tmem + i * BN,
                                                ptx::umma_desc_join(a_lo + (((j * MT + i) * C::A_TILE + 32 * k) >> 4), HI),
                                                ptx::umma_desc_join(b_lo + ((j * C::B_TILE + 32 * k) >> 4), HI), idesc,
                                                ((ks - ks_lo) | j | k) != 0 ? 1u : 0u);
                        }
                    }
                    ptx::umma_commit(bar_empty + 8 * stage);   // smem slot free once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(bar_tfull + 8 * acc);          // accumulator complete
                stamp(p, 4);
            }
        }
        if constexpr (kClusterCapable) {
            if (p.cluster_splitk && p.splits > 1 && (int)blockIdx.x < num_tiles) { __syncwarp(); ptx::cluster_sync_all(); ptx::cluster_sync_all(); }
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        const int quarter = warp & 3;                  // TMEM lanes [32*quarter, +32)
        const int row = quarter * 32 + lane;           // tile row = output pixel slot
        uint32_t red_phase = 0;
        int local = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
            const int acc = local % ACC;
            const uint32_t acc_phase = (local / ACC) & 1;
            const int n_tile = item_mn(tile) % p.num_n_tiles;
            const int m_tile = item_mn(tile) / p.num_n_tiles;
            ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
            ptx::tc_fence_after();
            if (row == 0) stamp(p, 5);

            const bool split_k = p.splits > 1;
            const int ldw = p.num_n_tiles * BN;
            if constexpr (kClusterCapable) {
                if (split_k && p.cluster_splitk) {
                    // ---- cluster split-K: park the partial tile in shared memory, reduce a column strip over the peers
                    constexpr uint32_t PITCH = (BN + 4) * 4;          // bytes per parked row: 16-byte stores stay conflict-free
                    static_assert(kBM * PITCH <= (uint32_t)STAGES * C::STAGE_BYTES, "the partial tile fits the idle ring");
                    const uint32_t my_row = smem_base + (uint32_t)row * PITCH;
#pragma unroll 1
                    for (int chunk = 0; chunk < BN / 32; ++chunk) {
                        uint32_t raw[32];
                        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * C::ACC_COLS + chunk * 32, raw);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + (chunk * 32 + j) * 4), "r"(raw[j]),
                                         "r"(raw[j + 1]), "r"(raw[j + 2]), "r"(raw[j + 3]) : "memory");
                    }
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(bar_tempty + 8 * acc);
                    ptx::cluster_sync_all();                          // #1: every peer's partial tile is parked (all warps of all CTAs arrive)
                    if (row == 0) stamp(p, 6);
                    const int rank = (int)ptx::cluster_ctarank();     // == item_split(tile)
                    const int cw = BN / p.splits;                     // strip width in floats: a multiple of 4 (splits <= 16 <= BN / 8)
                    // the strip goes through in sub-strips of at most 32 columns (the width of the pool staging area)
                    for (int sub = 0; sub < cw; sub += 32) {
                        const int sw = cw - sub < 32 ? cw - sub : 32;
                        const int c4n = sw >> 2;                          // 16-byte units per row of the sub-strip
                        // item = (row, 16-byte unit) of the strip, consecutive threads = consecutive units of a row: a warp's remote
                        // load is then a few whole 128-byte runs of one peer's shared memory (one scattered 16 bytes per lane ran
                        // at a tenth of the speed)
                        // A 2x2/2 pool never looks past the tile's own row pairs: the rows behind them are not fetched.
                        int n_rows = kBM;
                        if constexpr (POOL1) { if (p.pool_stride == 2) n_rows = p.m_stride; }
                        const int c4sh = __ffs(c4n) - 1;                  // BN and the cluster size are powers of two: so is c4n
                        for (int item = row; item < n_rows * c4n; item += 128) {
                            const int r = item >> c4sh, c4 = (item & (c4n - 1)) * 4;
                            const int m = p.m_base + m_tile * p.m_stride + r;
                            if (m >= p.M_total) break;                    // rows ascend with the item index
                            const uint32_t src = smem_base + (uint32_t)r * PITCH + (uint32_t)(rank * cw + sub + c4) * 4;
                            float4 t4[16];
#pragma unroll
                            for (int peer = 0; peer < 16; ++peer) {       // all peers' loads in flight together, then summed in rank order
                                if (peer < p.splits)
                                    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                                                 : "=f"(t4[peer].x), "=f"(t4[peer].y), "=f"(t4[peer].z), "=f"(t4[peer].w)
                                                 : "r"(ptx::mapa(src, (uint32_t)peer)));
                            }
                            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int peer = 0; peer < 16; ++peer) {
                                if (peer < p.splits) { sum.x += t4[peer].x; sum.y += t4[peer].y; sum.z += t4[peer].z; sum.w += t4[peer].w; }
                            }
                            float v4[4] = {sum.x, sum.y, sum.z, sum.w};
                            const int col = n_tile * BN + rank * cw + sub + c4;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float y = fmaf(v4[e], s_scale[col + e], s_shift[col + e]);
                                if (p.leaky) y = fmaxf(y, 0.1f * y);
                                if constexpr (OUT_BF16) y = __bfloat162float(__float2bfloat16_rn(y));
                                else if (p.round_tf32) y = ptx::round_tf32(y);
                                v4[e] = y;
                            }
                            if constexpr (POOL1) {
                                // the strip's rows all live in this CTA: stage the activated values, pool below
                                if (p.prepool != nullptr && r < p.m_stride) {
                                    for (int e = 0; e < 4; ++e)
                                        if (col + e < p.OC) p.prepool[(size_t)m * p.OC + col + e] = v4[e];
                                }
                                pool1::sts4(ptx::smem_u32(s_out) + (uint32_t)(r * pool1::kPitch + c4) * 4, v4[0], v4[1], v4[2], v4[3]);
                            } else if constexpr (OUT_BF16) {
                                __nv_bfloat16 *o = ptx::out_at<__nv_bfloat16>(p.out, (size_t)m * p.OC + col, max(0, min(4, p.OC - col)), p.out_elems, 7);
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (col + e < p.OC) o[e] = __float2bfloat16_rn(v4[e]);
                            } else {
                                float *o = ptx::out_at<float>(p.out, (size_t)m * p.OC + col, max(0, min(4, p.OC - col)), p.out_elems, 8);
                                if ((p.OC & 3) == 0 && col + 3 < p.OC) *reinterpret_cast<float4 *>(o) = make_float4(v4[0], v4[1], v4[2], v4[3]);
                                else {
#pragma unroll
                                    for (int e = 0; e < 4; ++e)
                                        if (col + e < p.OC) o[e] = v4[e];
                                }
                            }
                        }
                        if (row == 0) stamp(p, 9);
                        if constexpr (POOL1) {
                            asm volatile("bar.sync 1, 128;" ::: "memory");
                            const int m0 = p.m_base + m_tile * p.m_stride;
                            if (p.pool_stride == 2)
                                pool1::emit_s2<OUT_BF16>(ptx::smem_u32(s_out), quarter, lane, m0, p.m_stride, p.M_total, p.W, p.OC,
                                                         n_tile * BN + rank * cw + sub, sw, p.out, p.out_elems);
                            else
                                pool1::emit_s1<OUT_BF16>(ptx::smem_u32(s_out), quarter, lane, m0, p.m_stride, p.M_total, p.H, p.OC,
                                                         n_tile * BN + rank * cw + sub, sw, p.out, p.out_elems);
                            asm volatile("bar.sync 1, 128;" ::: "memory");       // the staging area is rewritten by the next sub-strip
                        }
                    }
                    if (row == 0) stamp(p, 7);
                    ptx::cluster_sync_all();                          // #2: nobody reads this CTA's shared memory any more
                    continue;
                }
            }
            if (split_k) {
                const int split = item_split(tile);
                const size_t slab = (size_t)p.num_m_tiles * MT * kBM * ldw;       // floats per K-slice slab
                // phase 1: this K slice's partial sums -> own slab; free the accumulator.  The tile is parked in the (idle)
                // shared-memory ring first and leaves as one bulk async copy per row: whole lines written by the copy
                // engine, instead of 64 scattered 16-byte stores per thread whose completion the fence below then had
                // to wait for (8 us of a 22 us kernel, tools/batch1_timeline.py).
                if constexpr (MT == 1) {
                    constexpr uint32_t PITCH = (BN + 4) * 4;
                    static_assert(kBM * PITCH <= (uint32_t)STAGES * C::STAGE_BYTES, "the partial tile fits the idle ring");
                    const uint32_t my_row = smem_base + (uint32_t)row * PITCH;
#pragma unroll 1
                    for (int chunk = 0; chunk < BN / 32; ++chunk) {
                        uint32_t raw[32];
                        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * C::ACC_COLS + chunk * 32, raw);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + (chunk * 32 + j) * 4), "r"(raw[j]),
                                         "r"(raw[j + 1]), "r"(raw[j + 2]), "r"(raw[j + 3]) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // parked by the generic proxy, read by the copy engine
                    if (p.m_base + m_tile * kBM + row < p.M_total) {       // rows past the last pixel are never read back
                        float *wrow = p.workspace + split * slab + ((size_t)m_tile * kBM + row) * ldw + n_tile * BN;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(wrow), "r"(my_row), "r"(BN * 4) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    if (row == 0) stamp(p, 8);
                } else {
#pragma unroll 1
                    for (int mi = 0; mi < MT; ++mi) {
                        float4 *wrow = reinterpret_cast<float4 *>(p.workspace + split * slab +
                                                                  ((size_t)(m_tile * MT + mi) * kBM + row) * ldw + n_tile * BN);
#pragma unroll 1
                        for (int chunk = 0; chunk < BN / 32; ++chunk) {
                            uint32_t raw[32];
                            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * C::ACC_COLS + mi * BN + chunk * 32, raw);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                __stcg(wrow + chunk * 8 + j, make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                                                          __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3])));
                        }
                    }
                }
                ptx::tc_fence_before();
                ptx::mbar_arrive(bar_tempty + 8 * acc);
                asm volatile("fence.proxy.async;" ::: "memory");      // the siblings read the slabs with bulk copies (async proxy)
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (row == 0) {
                    int *ctr = p.counters + 2 * item_mn(tile);
                    atomicAdd(ctr, 1);
                    long long t0 = clock64();
                    while (true) {                  // every sibling CTA is resident (grid <= SMs): this terminates
                        int seen;
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
                        if (seen >= p.splits) break;
                        __nanosleep(32);
                        if (clock64() - t0 > 4000000000LL) __trap();
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                __threadfence();
                asm volatile("fence.proxy.async;" ::: "memory");
                if (row == 0) stamp(p, 6);
                // phase 2: rows [r0, r1) of the tile are this CTA's share.  Their partial sums in every slab are
                // pulled into the (now idle) smem ring by bulk copies, one per (slab, row), all in flight at once
                // -- one L2 round trip instead of one per group of loads -- then summed IN SLAB ORDER, finished
                // and stored.
                const int tile_rows = MT * kBM;
                const int rows_per = (tile_rows + p.splits - 1) / p.splits;
                const int r0 = split * rows_per;
                int r1 = r0 + rows_per < tile_rows ? r0 + rows_per : tile_rows;
                if (m_tile * tile_rows + r1 > p.M_total) r1 = p.M_total - m_tile * tile_rows;     // rows past the last pixel
                const int nrows = r1 - r0;
                // Departure (the atomic's round trip is kept off the critical path: issued after the copies below
                // when there are any).  The last sibling to leave resets the pair for the next launch.
                auto depart = [&]() {
                    if (row == 0) {
                        int *ctr = p.counters + 2 * item_mn(tile);
                        if (atomicAdd(ctr + 1, 1) == p.splits - 1) {
                            ctr[1] = 0;
                            __threadfence();
                            ctr[0] = 0;
                        }
                    }
                };
                if (nrows <= 0) { depart(); continue; }
                constexpr uint32_t ROW_BYTES = BN * 4;
                if ((size_t)p.splits * nrows * ROW_BYTES > (size_t)STAGES * C::STAGE_BYTES) __trap();   // planner: MT = 1, splits <= 128
                if (row == 0) ptx::mbar_expect_tx(bar_red, (uint32_t)(p.splits * nrows) * ROW_BYTES);
                for (int c = row; c < p.splits * nrows; c += 128) {
                    const int sidx = c / nrows, r = c - sidx * nrows;
                    const float *src = p.workspace + sidx * slab + ((size_t)m_tile * tile_rows + r0 + r) * ldw + n_tile * BN;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_base + (uint32_t)c * ROW_BYTES), "l"(src), "r"(ROW_BYTES), "r"(bar_red) : "memory");
                }
                depart();
                ptx::mbar_wait(bar_red, red_phase);
                red_phase ^= 1;
                if (row == 0) stamp(p, 9);
                constexpr int C4 = BN / 4;
                const float4 *red = reinterpret_cast<const float4 *>(smem);
                for (int u = row; u < nrows * C4; u += 128) {
                    const int r = u / C4, c4 = u - r * C4;
                    const int m = m_tile * tile_rows + r0 + r;                 // output pixel (im2col: linear index)
                    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int sidx = 0; sidx < p.splits; ++sidx) {
                        const float4 t4 = red[(size_t)(sidx * nrows + r) * C4 + c4];
                        sum.x += t4.x; sum.y += t4.y; sum.z += t4.z; sum.w += t4.w;
                    }
                    const int col = n_tile * BN + c4 * 4;
                    float v4[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float y = fmaf(v4[e], s_scale[col + e], s_shift[col + e]);
                        if (p.leaky) y = fmaxf(y, 0.1f * y);
                        if constexpr (OUT_BF16) y = __bfloat162float(__float2bfloat16_rn(y));
                        else if (p.round_tf32) y = ptx::round_tf32(y);
                        v4[e] = y;
                    }
                    if constexpr (OUT_BF16) {
                        __nv_bfloat16 *o = ptx::out_at<__nv_bfloat16>(p.out, (size_t)m * p.OC + col, max(0, min(4, p.OC - col)), p.out_elems, 7);
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (col + e < p.OC) o[e] = __float2bfloat16_rn(v4[e]);
                    } else {
                        float *o = ptx::out_at<float>(p.out, (size_t)m * p.OC + col, max(0, min(4, p.OC - col)), p.out_elems, 8);
                        if ((p.OC & 3) == 0 && col + 3 < p.OC) *reinterpret_cast<float4 *>(o) = make_float4(v4[0], v4[1], v4[2], v4[3]);
                        else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (col + e < p.OC) o[e] = v4[e];
                        }
                    }
                }
                if (row == 0) stamp(p, 7);
                continue;
            }

#pragma unroll 1
            for (int mi = 0; mi < MT; ++mi) {
                bool valid;
                size_t out_pix;   // index of the stored pixel (pooled pixel when POOL)
                size_t pre_pix = 0;   // index of the conv pixel (debug pre-pool dump)
                if constexpr (IM2COL) {
                    const int m = p.m_base + (m_tile * MT + mi) * p.m_stride + row;
                    valid = m < p.M_total;
                    out_pix = (size_t)m;
                } else {
                    const int w0 = (m_tile % p.tiles_w) * p.tw;
                    const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.th;
                    const int n0 = ((m_tile / (p.tiles_w * p.tiles_h)) * MT + mi) * p.nb;
                    const int wl = row & (p.tw - 1);
                    const int hl = (row >> p.log2_tw) & (p.th - 1);
                    const int nl = row >> (p.log2_tw + p.log2_th);
                    const int n = n0 + nl, h = h0 + hl, w = w0 + wl;
                    valid = n < p.B;
                    if constexpr (POOL) out_pix = ((size_t)n * p.OH + (h >> 1)) * p.OW + (w >> 1);
                    else out_pix = ((size_t)n * p.OH + h) * p.OW + w;
                    pre_pix = ((size_t)n * p.H + h) * p.W + w;
                }

#pragma unroll 1
                for (int chunk = 0; chunk < BN / 32; ++chunk) {
                    uint32_t raw[32];
                    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * C::ACC_COLS + mi * BN + chunk * 32, raw);
                    ptx::tmem_ld_wait();
                    const int col0 = n_tile * BN + chunk * 32;
                    float v[32];
                    if constexpr (POOL_FIRST) {
                        // every scale > 0: scale/shift, LeakyReLU and rounding are monotone, so the window maximum of the raw
                        // accumulators is taken first and one value is activated instead of four (bit-identical)
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + col0 + j);
                            const float4 sh = *reinterpret_cast<const float4 *>(s_shift + col0 + j);
                            v[j + 0] = fmaf(__uint_as_float(raw[j + 0]), sc.x, sh.x);
                            v[j + 1] = fmaf(__uint_as_float(raw[j + 1]), sc.y, sh.y);
                            v[j + 2] = fmaf(__uint_as_float(raw[j + 2]), sc.z, sh.z);
                            v[j + 3] = fmaf(__uint_as_float(raw[j + 3]), sc.w, sh.w);
                        }
                        if (p.leaky) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.1f * v[j]);
                        }
                        if constexpr (!OUT_BF16) {
                            if (p.round_tf32) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = ptx::round_tf32(v[j]);
                            }
                        } else {   // round before pooling: the reference pools the stored values
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
                        }
                    }

                    if constexpr (POOL1) {
                        pool1::dump_prepool(v, (int)out_pix, p.M_total, row < p.m_stride, p.OC, col0, p.prepool);
                        if (p.pool_stride == 2)
                            pool1::store_chunk_s2<OUT_BF16>(ptx::smem_u32(s_out), v, quarter, lane, p.m_base + (m_tile * MT + mi) * p.m_stride,
                                                            p.m_stride, p.M_total, p.W, p.OC, col0, p.out, p.out_elems);
                        else
                            pool1::store_chunk<OUT_BF16>(ptx::smem_u32(s_out), v, quarter, lane, p.m_base + (m_tile * MT + mi) * p.m_stride,
                                                         p.m_stride, p.M_total, p.H, p.OC, col0, p.out, p.prepool, p.out_elems);
                    } else if constexpr (POOL) {
                        if constexpr (!POOL_FIRST) {      // debug dump of the pooled-over activations (the host selects this variant for it)
                            if (p.prepool != nullptr && valid) {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (col0 + j < p.OC) p.prepool[pre_pix * p.OC + col0 + j] = v[j];
                            }
                        }
                        // butterfly 2x2 max over lanes (w^1, then h^1 = lane ^ tw); each lane ends up
                        // with 8 of the 32 columns of its window: 16*(w&1) + 8*(h&1) + [0,8)
                        const bool wodd = lane & 1;
                        float a[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float keep = wodd ? v[16 + j] : v[j];
                            const float send = wodd ? v[j] : v[16 + j];
                            a[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
                        }
                        const bool hodd = (lane >> p.log2_tw) & 1;
                        float r8[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float keep = hodd ? a[8 + j] : a[j];
                            const float send = hodd ? a[j] : a[8 + j];
                            r8[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, p.tw));
                        }
                        const int col = col0 + (wodd ? 16 : 0) + (hodd ? 8 : 0);
                        if constexpr (POOL_FIRST) {
                            const float4 sca = *reinterpret_cast<const float4 *>(s_scale + col), scb = *reinterpret_cast<const float4 *>(s_scale + col + 4);
                            const float4 sha = *reinterpret_cast<const float4 *>(s_shift + col), shb = *reinterpret_cast<const float4 *>(s_shift + col + 4);
                            const float sc8[8] = {sca.x, sca.y, sca.z, sca.w, scb.x, scb.y, scb.z, scb.w};
                            const float sh8[8] = {sha.x, sha.y, sha.z, sha.w, shb.x, shb.y, shb.z, shb.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float y = fmaf(r8[j], sc8[j], sh8[j]);
                                if (p.leaky) y = fmaxf(y, 0.1f * y);
                                if constexpr (OUT_BF16) y = __bfloat162float(__float2bfloat16_rn(y));
                                else if (p.round_tf32) y = ptx::round_tf32(y);
                                r8[j] = y;
                            }
                        }
                        if (valid && col < p.OC) {
                            if constexpr (OUT_BF16) {
                                __nv_bfloat16 *o = ptx::out_at<__nv_bfloat16>(p.out, out_pix * p.OC + col, 8, p.out_elems, 9);
                                uint4 pk;
                                __nv_bfloat162 t0 = __floats2bfloat162_rn(r8[0], r8[1]);
                                __nv_bfloat162 t1 = __floats2bfloat162_rn(r8[2], r8[3]);
                                __nv_bfloat162 t2 = __floats2bfloat162_rn(r8[4], r8[5]);
                                __nv_bfloat162 t3 = __floats2bfloat162_rn(r8[6], r8[7]);
                                pk.x = *reinterpret_cast<uint32_t *>(&t0);
                                pk.y = *reinterpret_cast<uint32_t *>(&t1);
                                pk.z = *reinterpret_cast<uint32_t *>(&t2);
                                pk.w = *reinterpret_cast<uint32_t *>(&t3);
                                *reinterpret_cast<uint4 *>(o) = pk;
                            } else {
                                float *o = ptx::out_at<float>(p.out, out_pix * p.OC + col, 8, p.out_elems, 10);
                                *reinterpret_cast<float4 *>(o) = make_float4(r8[0], r8[1], r8[2], r8[3]);
                                *reinterpret_cast<float4 *>(o + 4) = make_float4(r8[4], r8[5], r8[6], r8[7]);
                            }
                        }
                    } else {
                        if (valid) {
                            if constexpr (OUT_BF16) {
                                __nv_bfloat16 *o = ptx::out_at<__nv_bfloat16>(p.out, out_pix * p.OC + col0, max(0, min(32, p.OC - col0)), p.out_elems, 11);
                                if ((p.OC & 7) == 0) {
#pragma unroll
                                    for (int j = 0; j < 32; j += 8) {
                                        if (col0 + j < p.OC) {
                                            uint4 pk;
                                            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]);
                                            __nv_bfloat162 t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                                            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
                                            __nv_bfloat162 t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                                            pk.x = *reinterpret_cast<uint32_t *>(&t0);
                                            pk.y = *reinterpret_cast<uint32_t *>(&t1);
                                            pk.z = *reinterpret_cast<uint32_t *>(&t2);
                                            pk.w = *reinterpret_cast<uint32_t *>(&t3);
                                            *reinterpret_cast<uint4 *>(o + j) = pk;
                                        }
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        if (col0 + j < p.OC) o[j] = __float2bfloat16_rn(v[j]);
                                }
                            } else if constexpr (STAGE_OUT) {
                                float *so = s_out + (size_t)(quarter * 32 + lane) * p.OC + col0;
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (col0 + j < p.OC) so[j] = v[j];
                            } else {
                                float *o = ptx::out_at<float>(p.out, out_pix * p.OC + col0, max(0, min(32, p.OC - col0)), p.out_elems, 12);
                                if ((p.OC & 3) == 0) {
#pragma unroll
                                    for (int j = 0; j < 32; j += 4)
                                        if (col0 + j < p.OC)
                                            *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        if (col0 + j < p.OC) o[j] = v[j];
                                }
                            }
                        }
                    }
                }
                if constexpr (STAGE_OUT) {
                    // rows are OC floats wide and adjacent in memory: this warp's 32 rows are one contiguous run, written by
                    // ONE bulk async copy shared -> global (cp.async.bulk: the copy engine streams whole lines, no per-lane
                    // store instructions) whenever both ends fall on 16-byte boundaries; by hand otherwise
                    const int row0 = p.m_base + m_tile * p.m_stride + quarter * 32;
                    const int rows = min(32, p.M_total - row0);
                    const float *src = s_out + (size_t)quarter * 32 * p.OC;
                    float *dst = ptx::out_at<float>(p.out, (size_t)row0 * p.OC, rows > 0 ? (size_t)rows * p.OC : 0, p.out_elems, 13);
                    const uint32_t bytes = rows > 0 ? (uint32_t)rows * p.OC * 4 : 0;
                    const bool bulk = p.tma_store && bytes > 0 && (bytes & 15) == 0 && (ptx::smem_u32(src) & 15) == 0 &&
                                      (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
                    if (bulk) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the staged rows were written by the generic proxy
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(ptx::smem_u32(src)), "r"(bytes) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // the staging rows may be rewritten
                        }
                        __syncwarp();
                    } else {
                        __syncwarp();
                        if (rows > 0) {
                            const int count = rows * p.OC;
                            if ((((size_t)quarter * 32 * p.OC) & 3) == 0 && (((size_t)row0 * p.OC) & 3) == 0) {
                                for (int i = lane * 4; i + 3 < count; i += 128)
                                    *reinterpret_cast<float4 *>(dst + i) = *reinterpret_cast<const float4 *>(src + i);
                                for (int i = (count & ~3) + lane; i < count; i += 32) dst[i] = src[i];
                            } else {
                                for (int i = lane; i < count; i += 32) dst[i] = src[i];
                            }
                        }
                        __syncwarp();
                    }
                }
            }   // mi
            if (row == 0) stamp(p, 7);
            ptx::tc_fence_before();
            ptx::mbar_arrive(bar_tempty + 8 * acc);
        }
        if constexpr (STAGE_OUT) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // bulk stores have reached global memory
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Tensor-pipe peak probe: every CTA issues `iters` x (128 x 256 x 32B-of-K) MMAs back to back
// on resident shared memory, no loads, no epilogue.  `random_data`: fill the operands with
// pseudo-random values in (-1, 1) instead of zeros -- the power a dense MMA draws, hence the clock
// the part sustains under its power cap, depends on the data toggling.
template <bool BF16>
__global__ void __launch_bounds__(64, 1) mma_peak_kernel(int iters, int random_data) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int A_BYTES = 128 * 128, B_BYTES = 256 * 128;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) {
        uint32_t word = 0;
        if (random_data) {
            uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u + 12345u;
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            if constexpr (BF16) {
                // two bf16 in (-1, 1): sign, exponent 0x7e (0.5 .. 1), random mantissa
                const uint32_t lo = ((h & 0x8000u) | 0x3f00u | (h & 0x7fu));
                const uint32_t hi = (((h >> 16) & 0x8000u) | 0x3f00u | ((h >> 16) & 0x7fu));
                word = lo | (hi << 16);
            } else {
                word = (h & 0x80000000u) | 0x3f000000u | (h & 0x007fe000u);   // tf32-exact fp32 in (-1, 1)
            }
        }
        reinterpret_cast<uint32_t *>(smem)[i] = word;
    }
    const uint32_t b = ptx::smem_u32(&bar);
    if (threadIdx.x == 0) { ptx::mbar_init(b, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(&slot), 512);
    // generic-proxy writes above must be visible to the async proxy (UMMA reads smem)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = ptx::umma_idesc(BF16 ? 1u : 2u, 256);
        const uint64_t da = ptx::umma_desc_kmajor(ptx::smem_u32(smem), 128);
        const uint64_t db = ptx::umma_desc_kmajor(ptx::smem_u32(smem) + A_BYTES, 128);
        uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma<BF16>(tmem + (it & 1) * 256, da + 2 * k, db + 2 * k, idesc, 1u);
            if ((it & 31) == 31 || it == iters - 1) {
                ptx::umma_commit(b);
                ptx::mbar_wait(b, phase);
                phase ^= 1;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

}  // namespace umma_conv
