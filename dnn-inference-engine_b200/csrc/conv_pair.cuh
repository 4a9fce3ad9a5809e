// conv_pair.cuh -- the long-K 13x13 blocks (5-7) on CTA PAIRS: tcgen05.mma.cta_group::2.
//
// Same implicit GEMM as conv_umma.cuh's im2col mode (M = output pixels, N = output channels,
// K = (tap, channel chunk); A tiles by TMA im2col straight from the NHWC activation, B tiles from the
// prepacked [tap][chunk][OCp][32] weights), but two CTAs on the two SMs of a TPC work on one
// 256 x 256 (MT = 1) or 512 x 256 (MT = 2) output tile together: each CTA loads its own 128 x MT rows
// of A and only HALF of the 256-row B tile; one thread of the leader CTA issues M = 256 MMAs that
// read both CTAs' shared memory and write both CTAs' TMEM.  Per output element that is a third
// (MT = 1: 32 KB instead of 48 KB per 128 x 256 x 32 block) or a quarter (MT = 2: 48 KB instead of
// 64 KB per 256 x 256 x 32) less L2 -> shared-memory traffic than the single-CTA kernel, and MT = 1
// keeps two accumulator stages and six smem stages.  Less traffic is what counts here: the whole
// step runs under the power cap, so bytes not moved are clock.
//
// Barriers (per stage s, accumulator stage a):
//   full[s]    leader's only; the leader's producer arms it with the bytes of BOTH CTAs, both CTAs'
//              TMA loads (cta_group::2 form) complete on it
//   empty[s]   one per CTA; tcgen05.commit multicast from the leader arrives on both
//   tfull[a]   one per CTA; multicast commit after the tile's last MMA
//   tempty[a]  leader's only; the 128 epilogue threads of each CTA arrive on it (remote arrive)
// Every K block is summed in the same order as in the single-CTA kernels, so results are identical.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>

#include "epilogue_pool1.cuh"
#include "ptx.cuh"

namespace pairconv {

constexpr int kThreads = 192;
constexpr int kBM = 128;
constexpr int kSwz = 128;

struct Params {
    int H, W, B, OC;
    int KH, KW, PT, PL;
    int c_chunks;             // 128-byte channel chunks per tap
    int num_m_tiles;          // pair tiles along M in this launch (2 * MT * 128 rows each)
    int num_n_tiles;
    int M_total;              // B * H * W
    int m_base;               // first output pixel of this launch
    int m_stride;             // output pixels each CTA's 128 rows advance: 128, or 128 - (W + 1) with the fused 2x2/1 pool
    float *prepool;           // POOL1 debug: also dump the pooled-over activations [B,H,W,OC] fp32
    const float *scale;
    const float *shift;
    void *out;
    int leaky, round_tf32;
    size_t out_elems;         // extent of `out` in elements (checked by -DDNNB_BOUNDS builds)
};

// kBN = output channels per pair tile (256; 128 for the tail launch over a block's last rows, so that
// they spread over twice as many pairs)
template <int MT, int kBN = 256, bool POOL1 = false>
struct Cfg {
    static constexpr int A_TILE = kBM * kSwz;
    static constexpr int B_HALF = (kBN / 2) * kSwz;
    static constexpr int STAGE_BYTES = MT * A_TILE + B_HALF;
    static constexpr int TAIL_BYTES = 1024 + 2 * 1024 * 4 + (POOL1 ? pool1::kStageBytes : 0);
    static constexpr int STAGES_RAW = (227 * 1024 - 1024 - TAIL_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + TAIL_BYTES + 1024;
    static constexpr int ACC_COLS = MT * kBN;
    static constexpr int ACC_STAGES = 2 * ACC_COLS <= 512 ? 2 : 1;
    static constexpr int TMEM_COLS = ACC_STAGES * ACC_COLS;
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS >= 32 && TMEM_COLS <= 512, "TMEM allocation");
};

template <bool BF16, bool OUT_BF16, int MT, int kBN = 256, bool POOL1 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p) {
    using C = Cfg<MT, kBN, POOL1>;
    constexpr int STAGES = C::STAGES;
    constexpr int ACC = C::ACC_STAGES;
    constexpr int ES = BF16 ? 2 : 4;
    constexpr int BKC = kSwz / ES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *tail = smem + STAGES * C::STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tail + 512);
    float *s_scale = reinterpret_cast<float *>(tail + 1024);
    float *s_shift = s_scale + 1024;
    float *s_stage = s_shift + 1024;      // POOL1: [128][33] staging of one column chunk

    const uint32_t smem_base = ptx::smem_u32(smem);
    const uint32_t bar_full = ptx::smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * STAGES;
    const uint32_t bar_tempty = bar_tfull + 16;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int num_kb = p.KH * p.KW * p.c_chunks;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(bar_full + 8 * s, 1);
            ptx::mbar_init(bar_empty + 8 * s, 1);
        }
        for (int s = 0; s < ACC; ++s) {
            ptx::mbar_init(bar_tfull + 8 * s, 1);
            ptx::mbar_init(bar_tempty + 8 * s, 2 * 128);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc_pair(ptx::smem_u32(tmem_slot), C::TMEM_COLS);
    for (int i = threadIdx.x; i < p.num_n_tiles * kBN; i += kThreads) {
        s_scale[i] = p.scale[i];
        s_shift[i] = p.shift[i];
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();          // both CTAs' barriers exist before anything remote touches them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch (no-ops when the launch does not ask for it): the next kernel's CTAs may take the SMs
    // this grid's ragged end leaves idle and run their prologue there; nothing above depends on the preceding kernel.
    ptx::griddep_launch();
    ptx::griddep_wait();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                const int n_tile = tile % p.num_n_tiles;
                const int m_tile = tile / p.num_n_tiles;
                int w0[MT], h0[MT], n0[MT];
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    const int m0 = p.m_base + ((m_tile * MT + i) * 2 + (int)rank) * p.m_stride;
                    n0[i] = m0 / (p.H * p.W);
                    const int rem = m0 - n0[i] * (p.H * p.W);
                    h0[i] = rem / p.W;
                    w0[i] = rem - h0[i] * p.W;
                }
                int tap = 0, cc = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t full = ptx::mapa(bar_full + 8 * stage, 0);
                    if (rank == 0) ptx::mbar_expect_tx(bar_full + 8 * stage, 2 * C::STAGE_BYTES);
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + MT * C::A_TILE;
                    const int di = tap / p.KW, dj = tap - di * p.KW;
#pragma unroll
                    for (int i = 0; i < MT; ++i)
                        ptx::tma_load_im2col_4d_pair(sa + i * C::A_TILE, &map_a, full, cc * BKC, w0[i] - p.PL, h0[i] - p.PT,
                                                     n0[i], (uint16_t)dj, (uint16_t)di);
                    ptx::tma_load_4d_pair(sb, &map_b, full, 0, n_tile * kBN + (int)rank * (kBN / 2), cc, tap);
                    if (++cc == p.c_chunks) { cc = 0; ++tap; }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0 && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::umma_idesc(BF16 ? 1u : 2u, kBN, 256);
            constexpr uint32_t HI = ptx::umma_desc_hi(kSwz);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
                const int acc = local % ACC;
                const uint32_t acc_phase = (local / ACC) & 1;
                ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(bar_full + 8 * stage, phase);
                    ptx::tc_fence_after();
                    const uint32_t a_lo = ptx::umma_desc_lo(smem_base + stage * C::STAGE_BYTES);
                    const uint32_t b_lo = a_lo + ((MT * C::A_TILE) >> 4);
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
#pragma unroll
                        for (int k = 0; k < kSwz / 32; ++k)
                            ptx::umma_pair<BF16>(d_tmem + i * kBN, ptx::umma_desc_join(a_lo + ((i * C::A_TILE + 32 * k) >> 4), HI),
                                                 ptx::umma_desc_join(b_lo + ((32 * k) >> 4), HI), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit_pair(bar_empty + 8 * stage, 3);     // both CTAs' smem slots are free once these retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit_pair(bar_tfull + 8 * acc, 3);           // accumulators complete in both CTAs
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (both CTAs, own TMEM)
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        int local = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
            const int acc = local % ACC;
            const uint32_t acc_phase = (local / ACC) & 1;
            const int n_tile = tile % p.num_n_tiles;
            const int m_tile = tile / p.num_n_tiles;
            ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int mi = 0; mi < MT; ++mi) {
                const int m_first = p.m_base + ((m_tile * MT + mi) * 2 + (int)rank) * p.m_stride;
                const int m = m_first + row;
                const bool valid = m < p.M_total;
                if constexpr (POOL1 && OUT_BF16) {
                    // bf16 storage: 64 columns per pass, pooled as packed pairs (epilogue_pool1.cuh: store_chunk_bf16x2)
                    auto activated = [&](int chunk, float (&v)[32]) {
                        uint32_t raw[32];
                        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * C::ACC_COLS + mi * kBN + chunk * 32, raw);
                        ptx::tmem_ld_wait();
                        const int c0 = n_tile * kBN + chunk * 32;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c0 + j);
                            const float4 sh = *reinterpret_cast<const float4 *>(s_shift + c0 + j);
                            v[j + 0] = fmaf(__uint_as_float(raw[j + 0]), sc.x, sh.x);
                            v[j + 1] = fmaf(__uint_as_float(raw[j + 1]), sc.y, sh.y);
                            v[j + 2] = fmaf(__uint_as_float(raw[j + 2]), sc.z, sh.z);
                            v[j + 3] = fmaf(__uint_as_float(raw[j + 3]), sc.w, sh.w);
                        }
                        if (p.leaky) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.1f * v[j]);
                        }
                        if (p.prepool != nullptr) {            // debug dump wants the stored (rounded) values
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
                            pool1::dump_prepool(v, m, p.M_total, row < p.m_stride, p.OC, c0, p.prepool);
                        }
                    };
#pragma unroll 1
                    for (int chunk = 0; chunk < kBN / 32; chunk += 2) {
                        float v0[32], v1[32];
                        activated(chunk, v0);
                        activated(chunk + 1, v1);
                        pool1::store_chunk_bf16x2(ptx::smem_u32(s_stage), v0, v1, quarter, lane, m_first, p.m_stride, p.M_total, p.H,
                                                  p.OC, n_tile * kBN + chunk * 32, p.out, p.out_elems);
                    }
                }
                constexpr int kChunks32 = (POOL1 && OUT_BF16) ? 0 : kBN / 32;       // (the packed path above took them all)
#pragma unroll 1
                for (int chunk = 0; chunk < kChunks32; ++chunk) {
                    uint32_t raw[32];
                    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * C::ACC_COLS + mi * kBN + chunk * 32, raw);
                    ptx::tmem_ld_wait();
                    const int col0 = n_tile * kBN + chunk * 32;
                    float v[32];
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
                    if constexpr (POOL1) {
                        if constexpr (OUT_BF16) {       // the reference pools the stored values: round first
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
                        } else if (p.round_tf32) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = ptx::round_tf32(v[j]);
                        }
                        pool1::dump_prepool(v, m, p.M_total, row < p.m_stride, p.OC, col0, p.prepool);
                        pool1::store_chunk<OUT_BF16>(ptx::smem_u32(s_stage), v, quarter, lane, m_first, p.m_stride, p.M_total, p.H,
                                                     p.OC, col0, p.out, p.prepool, p.out_elems);
                        continue;
                    }
                    if (!valid) continue;
                    if constexpr (OUT_BF16) {
                        __nv_bfloat16 *o = ptx::out_at<__nv_bfloat16>(p.out, (size_t)m * p.OC + col0, min(32, p.OC - col0), p.out_elems, 3);
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
                        if (p.round_tf32) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = ptx::round_tf32(v[j]);
                        }
                        float *o = ptx::out_at<float>(p.out, (size_t)m * p.OC + col0, min(32, p.OC - col0), p.out_elems, 4);
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (col0 + j < p.OC)
                                *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive_cluster(ptx::mapa(bar_tempty + 8 * acc, 0));     // the leader's barrier counts both CTAs
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();          // the peer may still be reading this CTA's smem / signalling its barriers
    if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
}

}  // namespace pairconv
