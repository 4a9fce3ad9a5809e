// launchers.cuh -- the kernel instantiations libdnn_b200.so carries and their dispatch tables:
// conv_umma_kernel (generic implicit GEMM), conv_pair_kernel (CTA pairs), conv_rowpair_kernel
// (patch-resident, blocks 1-2), plus launch_ex, the one place launch attributes are set.
#pragma once
#include <utility>

#include "conv_pair.cuh"
#include "conv_rowpair.cuh"
#include "conv_umma.cuh"
#include "host_util.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// Launch with attributes.  t_pdl is set around a forward of a latency-mode network: kernels are then
// launched with programmatic stream serialization, i.e. a kernel may start (prologue, weight prefetch)
// while its predecessor drains; each of these kernels calls ptx::griddep_wait() before it touches memory
// an earlier kernel of the stream writes.  Only kernels that contain that wait go through launch_ex.
// `coop` (split-K: sibling CTAs wait for one another) asks for a cooperative launch; under PDL the
// attribute is dropped -- the grid is at most one CTA per SM and a dependent kernel's CTAs are only
// scheduled after every CTA of this grid has started, so the siblings are co-resident either way.
thread_local bool t_pdl = false;

thread_local unsigned t_cluster = 0;    // > 1: the next launch_ex launches thread-block clusters of this many CTAs

template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool coop,
                      Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (t_cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = t_cluster;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
        coop = false;        // co-scheduling within a cluster is guaranteed by the hardware
    }
    if (t_pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    } else if (coop) {
        attr[n].id = cudaLaunchAttributeCooperative;
        attr[n].val.cooperative = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ------------------------------------------------------------------------------------------
// tcgen05 kernel instantiations and their dispatch table
typedef cudaError_t (*UmmaLauncher)(const CUtensorMap &, const CUtensorMap &, const umma_conv::Params &, int grid,
                                    cudaStream_t);
struct UmmaKey {
    bool bf16, out_bf16;
    int bn, swz, kbs;
    bool im2col, pool;
    int mt;   // 128-row M tiles per CTA tile
    bool stage_out;   // output tile staged in smem (rows that are not 16-byte multiples)
    bool pool1 = false;   // 2x2/1 SAME max-pool fused in the epilogue (im2col tiles overlapping by W + 1 rows)
};
typedef int (*UmmaClusterQuery)(int cluster_size);      // how many clusters of that size can be resident at once (0: not launchable)
struct UmmaEntry {
    UmmaKey key;
    UmmaLauncher launch;
    int stages, smem;
    UmmaClusterQuery max_clusters;
};

template <bool BF16, bool OUT_BF16, int BN, int SWZ, int KBS, bool IM2COL, bool POOL, int MT, bool SO = false, bool P1 = false>
int umma_max_clusters(int cluster_size) {
    using C = umma_conv::Cfg<BN, SWZ, KBS, MT, SO, P1>;
    auto kern = umma_conv::conv_umma_kernel<BF16, OUT_BF16, BN, SWZ, KBS, IM2COL, POOL, MT, SO, P1>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cluster_size > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)cluster_size * 16);
    cfg.blockDim = dim3(umma_conv::kThreads);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

template <bool BF16, bool OUT_BF16, int BN, int SWZ, int KBS, bool IM2COL, bool POOL, int MT, bool SO = false, bool P1 = false, bool PF = false>
cudaError_t launch_umma_t(const CUtensorMap &ma, const CUtensorMap &mb, const umma_conv::Params &p, int grid,
                          cudaStream_t st) {
    using C = umma_conv::Cfg<BN, SWZ, KBS, MT, SO, P1>;
    auto kern = umma_conv::conv_umma_kernel<BF16, OUT_BF16, BN, SWZ, KBS, IM2COL, POOL, MT, SO, P1, PF>;
    static bool done[64] = {};
    cudaError_t e = set_smem_once(kern, C::SMEM_BYTES, done);
    if (e != cudaSuccess) return e;
    if (p.cluster_splitk && p.splits > 1) {
        static bool np_done[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (p.splits > 8 && (dev < 0 || dev >= 64 || !np_done[dev])) {       // clusters of 16: beyond the portable size
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) np_done[dev] = true;
        }
        t_cluster = (unsigned)p.splits;
        e = launch_ex(kern, (unsigned)(grid < 0 ? -grid : grid), umma_conv::kThreads, C::SMEM_BYTES, st, false, ma, mb, p);
        t_cluster = 0;
        return e;
    }
    // grid < 0 = split-K: sibling CTAs wait for one another, so the whole grid must be resident
    return launch_ex(kern, (unsigned)(grid < 0 ? -grid : grid), umma_conv::kThreads, C::SMEM_BYTES, st, grid < 0, ma, mb, p);
}

template <bool BF16, bool OUT_BF16, int BN, int SWZ, int KBS, bool IM2COL, bool POOL, int MT, bool SO = false, bool P1 = false>
cudaError_t launch_umma(const CUtensorMap &ma, const CUtensorMap &mb, const umma_conv::Params &p, int grid,
                        cudaStream_t st) {
    if constexpr (POOL) {        // the fused 2x2/2 pool has a pool-first instantiation for all-positive scales
        if (p.pool_first && p.prepool == nullptr)
            return launch_umma_t<BF16, OUT_BF16, BN, SWZ, KBS, IM2COL, POOL, MT, SO, P1, true>(ma, mb, p, grid, st);
    }
    return launch_umma_t<BF16, OUT_BF16, BN, SWZ, KBS, IM2COL, POOL, MT, SO, P1, false>(ma, mb, p, grid, st);
}

#define UMMA_ENTRY_MT(BF16, OBF, BN, SWZ, KBS, I2C, POOL, MT)                                              \
    {{BF16, OBF, BN, SWZ, KBS, I2C, POOL, MT, false}, &launch_umma<BF16, OBF, BN, SWZ, KBS, I2C, POOL, MT>,  \
     umma_conv::Cfg<BN, SWZ, KBS, MT>::STAGES, umma_conv::Cfg<BN, SWZ, KBS, MT>::SMEM_BYTES,                 \
     &umma_max_clusters<BF16, OBF, BN, SWZ, KBS, I2C, POOL, MT>}
#define UMMA_ENTRY_SO(BF16, BN, SWZ, KBS)                                                                  \
    {{BF16, false, BN, SWZ, KBS, true, false, 1, true}, &launch_umma<BF16, false, BN, SWZ, KBS, true, false, 1, true>, \
     umma_conv::Cfg<BN, SWZ, KBS, 1, true>::STAGES, umma_conv::Cfg<BN, SWZ, KBS, 1, true>::SMEM_BYTES, nullptr}
#define UMMA_ENTRY_P1(BF16, OBF, BN)                                                                       \
    {{BF16, OBF, BN, 128, 1, true, false, 1, false, true}, &launch_umma<BF16, OBF, BN, 128, 1, true, false, 1, false, true>, \
     umma_conv::Cfg<BN, 128, 1, 1, false, true>::STAGES, umma_conv::Cfg<BN, 128, 1, 1, false, true>::SMEM_BYTES,  \
     &umma_max_clusters<BF16, OBF, BN, 128, 1, true, false, 1, false, true>}
#define UMMA_ENTRY(BF16, OBF, BN, SWZ, KBS, I2C, POOL) UMMA_ENTRY_MT(BF16, OBF, BN, SWZ, KBS, I2C, POOL, 1)

const UmmaEntry g_umma_table[] = {
    // TF32 operands (fp32 storage in and out)
    UMMA_ENTRY(false, false, 32, 64, 3, false, true),
    UMMA_ENTRY(false, false, 64, 128, 1, false, true),
    UMMA_ENTRY(false, false, 128, 128, 1, false, true),
    UMMA_ENTRY(false, false, 256, 128, 1, false, true),
    UMMA_ENTRY(false, false, 256, 128, 1, true, false),
    UMMA_ENTRY(false, false, 128, 128, 1, true, false),
    UMMA_ENTRY(false, false, 64, 128, 1, true, false),
    UMMA_ENTRY(false, false, 32, 128, 1, true, false),
    UMMA_ENTRY(false, false, 32, 64, 3, true, false),
    // BF16 operands, bf16 storage out
    UMMA_ENTRY(true, true, 32, 32, 3, false, true),
    UMMA_ENTRY(true, true, 64, 64, 3, false, true),
    UMMA_ENTRY(true, true, 128, 128, 1, false, true),
    UMMA_ENTRY(true, true, 256, 128, 1, false, true),
    UMMA_ENTRY(true, true, 256, 128, 1, true, false),
    UMMA_ENTRY(true, true, 128, 128, 1, true, false),
    UMMA_ENTRY(true, true, 64, 128, 1, true, false),
    UMMA_ENTRY(true, true, 32, 128, 1, true, false),
    UMMA_ENTRY(true, true, 64, 64, 3, true, false),
    UMMA_ENTRY(true, true, 32, 32, 3, true, false),
    // BF16 operands, fp32 out (a network's last block)
    UMMA_ENTRY(true, false, 32, 32, 3, false, true),
    UMMA_ENTRY(true, false, 64, 64, 3, false, true),
    UMMA_ENTRY(true, false, 128, 128, 1, false, true),
    UMMA_ENTRY(true, false, 256, 128, 1, false, true),
    UMMA_ENTRY(true, false, 256, 128, 1, true, false),
    UMMA_ENTRY(true, false, 128, 128, 1, true, false),
    UMMA_ENTRY(true, false, 64, 128, 1, true, false),
    UMMA_ENTRY(true, false, 32, 128, 1, true, false),
    UMMA_ENTRY(true, false, 64, 64, 3, true, false),
    UMMA_ENTRY(true, false, 32, 32, 3, true, false),
    // 256 x 256 CTA tiles (two M tiles share each weight tile) for the long-K 13x13 layers
    UMMA_ENTRY_MT(false, false, 256, 128, 1, true, false, 2),
    UMMA_ENTRY_MT(true, true, 256, 128, 1, true, false, 2),
    UMMA_ENTRY_MT(true, false, 256, 128, 1, true, false, 2),
    // the same for the fused-pool spatial tiles (two groups of nb images share each weight tile)
    UMMA_ENTRY_MT(false, false, 128, 128, 1, false, true, 2),
    UMMA_ENTRY_MT(false, false, 256, 128, 1, false, true, 2),
    UMMA_ENTRY_MT(true, true, 128, 128, 1, false, true, 2),
    UMMA_ENTRY_MT(true, true, 256, 128, 1, false, true, 2),
    // output staged through smem: the 125-channel head (block 8)
    UMMA_ENTRY_SO(false, 128, 128, 1),
    UMMA_ENTRY_SO(true, 128, 128, 1),
    // 2x2/1 SAME pool fused (block 5): whole block at small batches (BN 256), tail rows behind the CTA pairs (BN 128)
    UMMA_ENTRY_P1(false, false, 256), UMMA_ENTRY_P1(false, false, 128),
    UMMA_ENTRY_P1(true, true, 256),   UMMA_ENTRY_P1(true, true, 128),
    UMMA_ENTRY_P1(true, false, 256),  UMMA_ENTRY_P1(true, false, 128),
    // narrow N tiles for the latency plan: a short K loop that is not split spreads over the SMs by N instead
    UMMA_ENTRY_P1(false, false, 64),  UMMA_ENTRY_P1(false, false, 32),
    UMMA_ENTRY_P1(true, true, 64),    UMMA_ENTRY_P1(true, true, 32),
    UMMA_ENTRY_P1(true, false, 64),   UMMA_ENTRY_P1(true, false, 32),
};

const UmmaEntry *find_umma(const UmmaKey &k) {
    for (const UmmaEntry &e : g_umma_table)
        if (e.key.bf16 == k.bf16 && e.key.out_bf16 == k.out_bf16 && e.key.bn == k.bn && e.key.swz == k.swz &&
            e.key.kbs == k.kbs && e.key.im2col == k.im2col && e.key.pool == k.pool && e.key.mt == k.mt &&
            e.key.stage_out == k.stage_out && e.key.pool1 == k.pool1)
            return &e;
    return nullptr;
}

// ------------------------------------------------------------------------------------------
// CTA-pair kernel instantiations (conv_pair.cuh): tcgen05.mma.cta_group::2 for the long-K 13x13 blocks
typedef cudaError_t (*PairLauncher)(const CUtensorMap &, const CUtensorMap &, const pairconv::Params &, int grid,
                                    cudaStream_t);
struct PairEntry {
    bool bf16, out_bf16;
    int mt, bn;
    PairLauncher launch;
    int stages;
    bool pool1;
};
template <bool BF16, bool OUT_BF16, int MT, int BN, bool P1 = false>
cudaError_t launch_pair(const CUtensorMap &ma, const CUtensorMap &mb, const pairconv::Params &p, int grid, cudaStream_t st) {
    using C = pairconv::Cfg<MT, BN, P1>;
    auto kern = pairconv::conv_pair_kernel<BF16, OUT_BF16, MT, BN, P1>;
    static bool done[64] = {};
    cudaError_t e = set_smem_once(kern, C::SMEM_BYTES, done);
    if (e != cudaSuccess) return e;
    return launch_ex(kern, (unsigned)grid, pairconv::kThreads, C::SMEM_BYTES, st, false, ma, mb, p);     // __cluster_dims__(2,1,1): grid is even
}
#define PAIR_ENTRY(BF16, OBF, MT, BN) {BF16, OBF, MT, BN, &launch_pair<BF16, OBF, MT, BN>, pairconv::Cfg<MT, BN>::STAGES, false}
#define PAIR_ENTRY_P1(BF16, OBF, BN) {BF16, OBF, 1, BN, &launch_pair<BF16, OBF, 1, BN, true>, pairconv::Cfg<1, BN, true>::STAGES, true}
const PairEntry g_pair_table[] = {
    PAIR_ENTRY(false, false, 1, 256), PAIR_ENTRY(false, false, 2, 256), PAIR_ENTRY(false, false, 1, 128),
    PAIR_ENTRY(true, true, 1, 256),   PAIR_ENTRY(true, true, 2, 256),   PAIR_ENTRY(true, true, 1, 128),
    PAIR_ENTRY(true, false, 1, 256),  PAIR_ENTRY(true, false, 2, 256),  PAIR_ENTRY(true, false, 1, 128),
    // fused 2x2/1 SAME pool (block 5)
    PAIR_ENTRY_P1(false, false, 256), PAIR_ENTRY_P1(false, false, 128),
    PAIR_ENTRY_P1(true, true, 256),   PAIR_ENTRY_P1(true, true, 128),
    PAIR_ENTRY_P1(true, false, 256),  PAIR_ENTRY_P1(true, false, 128),
};
const PairEntry *find_pair(bool bf16, bool out_bf16, int mt, int bn = 256, bool pool1 = false) {
    for (const PairEntry &e : g_pair_table)
        if (e.bf16 == bf16 && e.out_bf16 == out_bf16 && e.mt == mt && e.bn == bn && e.pool1 == pool1) return &e;
    return nullptr;
}

// ------------------------------------------------------------------------------------------
// patch-resident row-pair kernel instantiations (conv_rowpair.cuh)
typedef cudaError_t (*RowpairLauncher)(const CUtensorMap &, const CUtensorMap &, const rowpair::Params &,
                                       const rowpair::LinMaps &, int grid, cudaStream_t);
struct RowpairEntry {
    bool bf16, out_bf16;
    int oc, swz, tr;        // tr = output rows per tile
    RowpairLauncher launch;         // 128-column tiles of the image
    RowpairLauncher launch_linear;  // tiles over the linearised (row group, column) positions
    int stages;
};

template <bool BF16, bool OUT_BF16, int OC, int SWZ, int TR, bool LINEAR>
cudaError_t launch_rowpair(const CUtensorMap &ma, const CUtensorMap &mw, const rowpair::Params &p,
                           const rowpair::LinMaps &lin, int grid, cudaStream_t st) {
    using C = rowpair::Cfg<OC, SWZ, TR, LINEAR>;
    auto kern = rowpair::conv_rowpair_kernel<BF16, OUT_BF16, OC, SWZ, TR, LINEAR>;
    static bool done[64] = {};
    cudaError_t e = set_smem_once(kern, C::SMEM_BYTES, done);
    if (e != cudaSuccess) return e;
    return launch_ex(kern, (unsigned)grid, rowpair::kThreads, C::SMEM_BYTES, st, false, ma, mw, p, lin);
}
#define ROWPAIR_ENTRY(BF16, OBF, OC, SWZ, TR)                                                             \
    {BF16, OBF, OC, SWZ, TR, &launch_rowpair<BF16, OBF, OC, SWZ, TR, false>, &launch_rowpair<BF16, OBF, OC, SWZ, TR, true>, \
     rowpair::Cfg<OC, SWZ, TR>::STAGES}
// per shape in order of preference (measured on B200, batch 512): 4-row tiles win for 16->32 channels
// (TF32 0.487 -> 0.468 ms, BF16 0.454 -> 0.396 ms), 2 rows for 32->64.  Taller tiles need fewer MMAs per pixel (-17 /
// -25 %) but leave two patch stages and two accumulator slots; measured in round 2 with the final kernels: 16->32 with
// 8 rows TF32 0.443 -> 0.489 ms, BF16 0.295 -> 0.309; 32->64 BF16 with 4 rows 0.159 -> 0.204.  Not instantiated;
// the planner takes the first entry whose TR divides the image height
const RowpairEntry g_rowpair_table[] = {
    ROWPAIR_ENTRY(false, false, 32, 64, 4), ROWPAIR_ENTRY(false, false, 32, 64, 2),
    ROWPAIR_ENTRY(false, false, 64, 128, 2),
    ROWPAIR_ENTRY(true, true, 32, 32, 4),   ROWPAIR_ENTRY(true, true, 32, 32, 2),
    ROWPAIR_ENTRY(true, true, 64, 64, 2),
    ROWPAIR_ENTRY(true, false, 32, 32, 4),  ROWPAIR_ENTRY(true, false, 32, 32, 2),
    ROWPAIR_ENTRY(true, false, 64, 64, 2),
};
const RowpairEntry *find_rowpair(bool bf16, bool out_bf16, int oc, int swz, int height, int max_tr) {
    for (const RowpairEntry &e : g_rowpair_table)
        if (e.bf16 == bf16 && e.out_bf16 == out_bf16 && e.oc == oc && e.swz == swz && height % e.tr == 0 && e.tr <= max_tr)
            return &e;
    return nullptr;
}
}  // namespace
