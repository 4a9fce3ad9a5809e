// engine.cu -- libdnn_b200.so: the C ABI declared in include/dnn_b200.h.
//   * op-level entry points with the reference's signatures (proj3/dnn_openblas.c)
//   * the resident network: per-block kernel planning, weight prepack, TMA descriptors,
//     HBM-resident activations, the forward launch sequence.
// Host-side only; the kernels live in conv_umma.cuh (tcgen05/TMA) and ops_simt.cuh.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/dnn_b200.h"
#include "conv0_umma.cuh"
#include "conv_pair.cuh"
#include "conv_rowpair.cuh"
#include "conv_umma.cuh"
#include "ops_simt.cuh"
#include "postproc.cuh"

namespace {

thread_local std::string g_err;

int fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CK(...)                                                                                       \
    do {                                                                                              \
        cudaError_t e_ = (__VA_ARGS__);                                                                    \
        if (e_ != cudaSuccess)                                                                        \
            return fail("%s:%d: %s -> %s", __FILE__, __LINE__, #__VA_ARGS__, cudaGetErrorString(e_));        \
    } while (0)

// ------------------------------------------------------------------------------------------
// driver entry points for tensor-map encoding (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const int *, const int *, cuuint32_t, cuuint32_t,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;

int load_driver_fns() {
    if (g_encode_tiled && g_encode_im2col) return 0;
    cudaDriverEntryPointQueryResult q;
    void *f = nullptr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    if (!f || q != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled not available");
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(f);
    f = nullptr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q));
    if (!f || q != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeIm2col not available");
    g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(f);
    return 0;
}

CUtensorMapSwizzle swizzle_enum(int swz) {
    return swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

// ------------------------------------------------------------------------------------------
// host-side rounding used by the weight prepack (round to nearest even)
float host_round_tf32(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u += 0xfffu + ((u >> 13) & 1u);
    u &= 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}
uint16_t host_bf16(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

int round_up(int a, int b) { return (a + b - 1) / b * b; }
int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

void out_pads(int in_size, int k, int s, int same, int *o, int *pf, int *pb) {
    // TF SAME / VALID, proj3/dnn_openblas.py:127-142
    if (same) {
        *o = (in_size + s - 1) / s;
        int pad = (*o - 1) * s + k - in_size;
        if (pad < 0) pad = 0;
        *pf = pad / 2;
        *pb = pad - pad / 2;
    } else {
        *o = (in_size - k + 1 + s - 1) / s;
        *pf = *pb = 0;
    }
}

// opt a kernel in to > 48 KB of dynamic shared memory, once per (kernel, device)
template <typename K>
cudaError_t set_smem_once(K kern, int bytes, bool (&done)[64]) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------
// Launch with attributes.  t_pdl is set around a forward of a latency-mode network: kernels are then
// launched with programmatic stream serialization, i.e. a kernel may start (prologue, weight prefetch)
// while its predecessor drains; each of these kernels calls ptx::griddep_wait() before it touches memory
// an earlier kernel of the stream writes.  Only kernels that contain that wait go through launch_ex.
// `coop` (split-K: sibling CTAs wait for one another) asks for a cooperative launch; under PDL the
// attribute is dropped -- the grid is at most one CTA per SM and a dependent kernel's CTAs are only
// scheduled after every CTA of this grid has started, so the siblings are co-resident either way.
thread_local bool t_pdl = false;

template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool coop,
                      Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    unsigned n = 0;
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
};
struct UmmaEntry {
    UmmaKey key;
    UmmaLauncher launch;
    int stages, smem;
};

template <bool BF16, bool OUT_BF16, int BN, int SWZ, int KBS, bool IM2COL, bool POOL, int MT, bool SO = false>
cudaError_t launch_umma(const CUtensorMap &ma, const CUtensorMap &mb, const umma_conv::Params &p, int grid,
                        cudaStream_t st) {
    using C = umma_conv::Cfg<BN, SWZ, KBS, MT, SO>;
    auto kern = umma_conv::conv_umma_kernel<BF16, OUT_BF16, BN, SWZ, KBS, IM2COL, POOL, MT, SO>;
    static bool done[64] = {};
    cudaError_t e = set_smem_once(kern, C::SMEM_BYTES, done);
    if (e != cudaSuccess) return e;
    // grid < 0 = split-K: sibling CTAs wait for one another, so the whole grid must be resident
    return launch_ex(kern, (unsigned)(grid < 0 ? -grid : grid), umma_conv::kThreads, C::SMEM_BYTES, st, grid < 0, ma, mb, p);
}

#define UMMA_ENTRY_MT(BF16, OBF, BN, SWZ, KBS, I2C, POOL, MT)                                              \
    {{BF16, OBF, BN, SWZ, KBS, I2C, POOL, MT, false}, &launch_umma<BF16, OBF, BN, SWZ, KBS, I2C, POOL, MT>,  \
     umma_conv::Cfg<BN, SWZ, KBS, MT>::STAGES, umma_conv::Cfg<BN, SWZ, KBS, MT>::SMEM_BYTES}
#define UMMA_ENTRY_SO(BF16, BN, SWZ, KBS)                                                                  \
    {{BF16, false, BN, SWZ, KBS, true, false, 1, true}, &launch_umma<BF16, false, BN, SWZ, KBS, true, false, 1, true>, \
     umma_conv::Cfg<BN, SWZ, KBS, 1, true>::STAGES, umma_conv::Cfg<BN, SWZ, KBS, 1, true>::SMEM_BYTES}
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
    UMMA_ENTRY(false, false, 32, 64, 3, true, false),
    // BF16 operands, bf16 storage out
    UMMA_ENTRY(true, true, 32, 32, 3, false, true),
    UMMA_ENTRY(true, true, 64, 64, 3, false, true),
    UMMA_ENTRY(true, true, 128, 128, 1, false, true),
    UMMA_ENTRY(true, true, 256, 128, 1, false, true),
    UMMA_ENTRY(true, true, 256, 128, 1, true, false),
    UMMA_ENTRY(true, true, 128, 128, 1, true, false),
    UMMA_ENTRY(true, true, 64, 64, 3, true, false),
    UMMA_ENTRY(true, true, 32, 32, 3, true, false),
    // BF16 operands, fp32 out (a network's last block)
    UMMA_ENTRY(true, false, 32, 32, 3, false, true),
    UMMA_ENTRY(true, false, 64, 64, 3, false, true),
    UMMA_ENTRY(true, false, 128, 128, 1, false, true),
    UMMA_ENTRY(true, false, 256, 128, 1, false, true),
    UMMA_ENTRY(true, false, 256, 128, 1, true, false),
    UMMA_ENTRY(true, false, 128, 128, 1, true, false),
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
};

const UmmaEntry *find_umma(const UmmaKey &k) {
    for (const UmmaEntry &e : g_umma_table)
        if (e.key.bf16 == k.bf16 && e.key.out_bf16 == k.out_bf16 && e.key.bn == k.bn && e.key.swz == k.swz &&
            e.key.kbs == k.kbs && e.key.im2col == k.im2col && e.key.pool == k.pool && e.key.mt == k.mt &&
            e.key.stage_out == k.stage_out)
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
};
template <bool BF16, bool OUT_BF16, int MT, int BN>
cudaError_t launch_pair(const CUtensorMap &ma, const CUtensorMap &mb, const pairconv::Params &p, int grid, cudaStream_t st) {
    using C = pairconv::Cfg<MT, BN>;
    auto kern = pairconv::conv_pair_kernel<BF16, OUT_BF16, MT, BN>;
    static bool done[64] = {};
    cudaError_t e = set_smem_once(kern, C::SMEM_BYTES, done);
    if (e != cudaSuccess) return e;
    kern<<<grid, pairconv::kThreads, C::SMEM_BYTES, st>>>(ma, mb, p);     // __cluster_dims__(2,1,1): grid is even
    return cudaGetLastError();
}
#define PAIR_ENTRY(BF16, OBF, MT, BN) {BF16, OBF, MT, BN, &launch_pair<BF16, OBF, MT, BN>, pairconv::Cfg<MT, BN>::STAGES}
const PairEntry g_pair_table[] = {
    PAIR_ENTRY(false, false, 1, 256), PAIR_ENTRY(false, false, 2, 256), PAIR_ENTRY(false, false, 1, 128),
    PAIR_ENTRY(true, true, 1, 256),   PAIR_ENTRY(true, true, 2, 256),   PAIR_ENTRY(true, true, 1, 128),
    PAIR_ENTRY(true, false, 1, 256),  PAIR_ENTRY(true, false, 2, 256),  PAIR_ENTRY(true, false, 1, 128),
};
const PairEntry *find_pair(bool bf16, bool out_bf16, int mt, int bn = 256) {
    for (const PairEntry &e : g_pair_table)
        if (e.bf16 == bf16 && e.out_bf16 == out_bf16 && e.mt == mt && e.bn == bn) return &e;
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
// (TF32 0.487 -> 0.468 ms, BF16 0.454 -> 0.396 ms; 8 rows leave too few patch stages), 2 rows for 32->64;
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

// ------------------------------------------------------------------------------------------
// Planner switches from the environment, read once per dnnb_net_finalize (DESIGN.md section 4c lists
// them).  Apart from DNNB_LATENCY_MODE they exist for bring-up, ablation measurements and tests that
// force a particular kernel; the defaults are the measured best plan.
struct Knobs {
    bool latency = false;            // DNNB_LATENCY_MODE=1   small-batch plan: split-K, PDL chain, graph replay
    bool no_pdl = false;             // DNNB_NO_PDL           latency plan without programmatic dependent launch
    bool no_graph = false;           // DNNB_NO_GRAPH         latency plan without CUDA-graph replay
    bool force_simt = false;         // DNNB_FORCE_SIMT       CUDA-core kernels everywhere
    bool no_conv0 = false;           // DNNB_NO_CONV0         block 0 through the generic kernel (after a cast)
    bool no_rowpair = false;         // DNNB_NO_ROWPAIR       blocks 1-2 through the generic kernel
    int rowpair_linear = -1;         // DNNB_ROWPAIR_LINEAR=0|1 force / DNNB_NO_ROWPAIR_LINEAR: linearised tiles off
    int rowpair_max_tr = 8;          // DNNB_ROWPAIR_MAX_TR   cap on the row-pair tile height
    bool no_pool_first = false;      // DNNB_NO_POOL_FIRST    always activate before pooling
    bool no_stage_out = false;       // DNNB_NO_STAGE_OUT     no smem staging of narrow odd output rows
    bool no_mt2 = false;             // DNNB_NO_MT2           no 256-row CTA tiles
    int mt2_mask = -1;               // DNNB_MT2_MASK         bit i: 256-row CTA tiles for block i
    bool no_pair = false;            // DNNB_NO_PAIR          no CTA-pair kernel
    int pair_mt = 1;                 // DNNB_PAIR_MT=2        512 x 256 pair tiles
    int pair_min_tiles = -1;         // DNNB_PAIR_MIN_TILES   pair tiles needed to use the pair kernel (default: two waves)
    bool pair_no_tail = false;       // DNNB_PAIR_NO_TAIL     pairs also run the partial last wave
    bool no_pair_tail = false;       // DNNB_NO_PAIR_TAIL     rows past the full waves go to the single-CTA kernel only
    bool no_tail = false;            // DNNB_NO_TAIL          no main/tail cut of wave-quantised 256 x 256 blocks
    int lat_fuse_nb = 0;             // DNNB_LAT_FUSE_NB      latency plan: keep fused spatial tiles up to this many images
    int split_min_stages = 0;        // DNNB_SPLIT_MIN_STAGES latency plan: K blocks per split (0: default 4, short loops whole)
    bool no_u8_fusion = false;       // DNNB_NO_U8_FUSION     uint8 frames through a separate preprocessing kernel
};

Knobs read_knobs() {
    auto flag = [](const char *name) { return getenv(name) != nullptr; };
    auto num = [](const char *name, int dflt) { const char *v = getenv(name); return v ? atoi(v) : dflt; };
    Knobs k;
    const char *lat = getenv("DNNB_LATENCY_MODE");
    k.latency = lat && lat[0] == '1';
    k.no_pdl = flag("DNNB_NO_PDL");
    k.no_graph = flag("DNNB_NO_GRAPH");
    k.force_simt = flag("DNNB_FORCE_SIMT");
    k.no_conv0 = flag("DNNB_NO_CONV0");
    k.no_rowpair = flag("DNNB_NO_ROWPAIR");
    k.rowpair_linear = flag("DNNB_ROWPAIR_LINEAR") ? (num("DNNB_ROWPAIR_LINEAR", 0) != 0) : flag("DNNB_NO_ROWPAIR_LINEAR") ? 0 : -1;
    k.rowpair_max_tr = num("DNNB_ROWPAIR_MAX_TR", 8);
    k.no_pool_first = flag("DNNB_NO_POOL_FIRST");
    k.no_stage_out = flag("DNNB_NO_STAGE_OUT");
    k.no_mt2 = flag("DNNB_NO_MT2");
    k.mt2_mask = num("DNNB_MT2_MASK", -1);
    k.no_pair = flag("DNNB_NO_PAIR");
    k.pair_mt = num("DNNB_PAIR_MT", 1) == 2 ? 2 : 1;
    k.pair_min_tiles = num("DNNB_PAIR_MIN_TILES", -1);
    k.pair_no_tail = flag("DNNB_PAIR_NO_TAIL");
    k.no_pair_tail = flag("DNNB_NO_PAIR_TAIL");
    k.no_tail = flag("DNNB_NO_TAIL");
    k.lat_fuse_nb = num("DNNB_LAT_FUSE_NB", 0);
    k.split_min_stages = num("DNNB_SPLIT_MIN_STAGES", 0);
    k.no_u8_fusion = flag("DNNB_NO_U8_FUSION");
    return k;
}

// ------------------------------------------------------------------------------------------
struct Block {
    // description
    int kh = 0, kw = 0, ic = 0, oc = 0, same = 1;
    bool has_bn = false, leaky = false;
    int pool_k = 0, pool_s = 0, pool_same = 1;
    std::vector<float> kernel, scale, shift;   // HWIO fp32; folded epilogue constants [oc]
    // geometry
    int ih = 0, iw = 0;                         // input spatial
    int ch = 0, cw = 0, pt = 0, pl = 0;         // conv output spatial, top/left pad
    int oh = 0, ow = 0, ppt = 0, ppl = 0;       // block output spatial (after pool), pool pads
    // plan
    int family = DNNB_KERNEL_SIMT;
    bool pool_fused = false, pool_separate = false;
    bool in_bf16 = false, out_bf16 = false;     // storage types
    bool round_tf32 = false;
    const UmmaEntry *umma = nullptr;
    bool op_bf16 = false;                       // conv0: MMA operand type (its input is always fp32)
    conv0::Params c0{};
    const RowpairEntry *rowpair = nullptr;
    rowpair::Params rp{};
    umma_conv::Params up{};
    CUtensorMap map_a{}, map_b{};
    // tail launch (im2col blocks whose tile count is not a whole number of waves): the rows left after
    // the last full wave, in finer tiles (smaller BN, MT = 1) so that they spread over all SMs
    bool rp_linear = false;              // row-pair kernel: tiles over linearised positions (conv_rowpair.cuh LINEAR)
    rowpair::LinMaps lin_maps{};
    const PairEntry *pair = nullptr;     // main launch on CTA pairs (cta_group::2) instead of b.umma
    pairconv::Params pp{};
    CUtensorMap map_b_pair{};            // weights map with a 128-row box: each CTA of a pair loads half a B tile
    // the last whole pair tiles of a block, past the full waves: pairs again, but 256 x 128 tiles (64-row B halves)
    const PairEntry *pair_tail = nullptr;
    pairconv::Params pp_tail{};
    CUtensorMap map_b_pair_tail{};
    int grid_pair_tail = 0;
    const UmmaEntry *umma_tail = nullptr;
    CUtensorMap map_b_tail{};
    umma_conv::Params up_tail{};
    int grid_tail = 0;
    const void *map_a_base = nullptr;           // input pointer the A map was encoded for
    int bn = 0, swz = 0, grid = 0;
    // device memory
    float *d_kernel = nullptr;                  // HWIO fp32 (SIMT)
    void *d_wpack = nullptr;                    // [tap][OCp][IC] tf32-rounded fp32 or bf16
    float *d_scale = nullptr, *d_shift = nullptr;
    float *d_workspace = nullptr;               // split-K partial sums, one slab per K slice
    int *d_counters = nullptr;                  // split-K arrival / departure counters per tile (zero between launches)
    void *d_conv_out = nullptr;                 // conv(+fused pool) output
    void *d_out = nullptr;                      // block output (== d_conv_out unless pool_separate)
    size_t conv_out_elems = 0, out_elems = 0;
};

}  // namespace

struct dnnb_net {
    int device = 0, precision = DNNB_TF32;
    int B = 0, H = 0, W = 0, C = 0;
    std::vector<Block> blocks;
    bool finalized = false;
    cudaStream_t stream = nullptr;
    float *d_in = nullptr;        // fp32 staging of the host input (= d_in2[0])
    // host path pipelining: H2D of forward i+1 and D2H of forward i-1 overlap the kernels of forward i
    float *d_in2[2] = {nullptr, nullptr};
    float *d_outf[2] = {nullptr, nullptr};   // final fp32 outputs, alternating (d_outf[0] is the last block's own buffer)
    const float *last_final = nullptr;       // where the most recent forward left its result
    const void *last_in = nullptr;           // fp32 device input of the most recent forward
    // uint8 frames (proj3/__init__.py:8-12 without the resize): staging + the /255 table
    uint8_t *d_u8[2] = {nullptr, nullptr};
    size_t u8_capacity = 0;       // bytes per staging buffer
    float *d_lut = nullptr;
    uint32_t *d_lut_conv0 = nullptr;   // the same table as finished block-0 operand bits (conv0_umma.cuh IN_U8)
    bool cur_in_u8 = false;            // set by forward_impl for the block-0 launch
    int cur_swap_rb = 0;
    // cv2-exact bilinear resize of frames that are not in_h x in_w: per-axis offset / weight tables
    int *d_resize_tab = nullptr;  // [xofs W][xw W][yofs H][yw H]
    int resize_sh = 0, resize_sw = 0;
    // on-device detection decoding (postproc.cuh): label boxes + counts, double-buffered like the outputs
    void *d_dets[2] = {nullptr, nullptr};
    int *d_counts[2] = {nullptr, nullptr};
    float score_thr = 0.3f;      // proj3/yolov2tiny.py:156
    double iou_thr = 0.3;        // :164
    int max_det = 32;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t in_ready[2] = {}, in_free[2] = {}, out_ready[2] = {}, out_free[2] = {};
    int flip = 0;
    void *d_in_cast = nullptr;    // input in the first tcgen05 block's storage type
    void *d_flush = nullptr;      // L2 flush scratch for dnnb_net_time_blocks
    int num_sms = 148;
    int launches = 0;
    // optional in-situ timing of one block (bench.py's roofline leg): event pairs recorded
    // around that block's kernels on whatever stream the forward runs on
    int prof_block = -1;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    // latency path (DNNB_LATENCY_MODE=1 at finalize): programmatic dependent launches, and the launch
    // sequence of dnnb_net_forward_device replayed from a CUDA graph per (input, output) pointer pair
    Knobs knobs;
    bool pdl = false, use_graph = false;
    struct GraphSlot {
        const void *in = nullptr;
        void *out = nullptr;
        cudaGraphExec_t exec = nullptr;
        unsigned long long used = 0;
    };
    std::vector<GraphSlot> graphs;
    unsigned long long graph_tick = 0;
    int block_limit = 0;                // debug: forwards stop after this many blocks (0 = all)
    const void *warm_in = nullptr;      // pointers of the last directly launched forward: the next forward on the
    void *warm_out = nullptr;           // same pair is captured (function attributes and tensor maps are set by then)
};

namespace {

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

int encode_map_a(dnnb_net *net, Block &b, const void *in_ptr) {
    const int es = b.in_bf16 ? 2 : 4;
    const CUtensorMapDataType dt = b.in_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    cuuint64_t dims[4] = {(cuuint64_t)b.ic, (cuuint64_t)b.iw, (cuuint64_t)b.ih, (cuuint64_t)net->B};
    cuuint64_t strides[3] = {(cuuint64_t)b.ic * es, (cuuint64_t)b.iw * b.ic * es, (cuuint64_t)b.ih * b.iw * b.ic * es};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const int bkc = b.swz / es;
    CUresult r;
    if (b.family == DNNB_KERNEL_UMMA_IM2COL) {
        int lower[2] = {-b.pl, -b.pt};
        int upper[2] = {(b.kw - 1 - b.pl) - (b.kw - 1), (b.kh - 1 - b.pt) - (b.kh - 1)};   // pad_back - (k-1)
        r = g_encode_im2col(&b.map_a, dt, 4, const_cast<void *>(in_ptr), dims, strides, lower, upper, (cuuint32_t)bkc,
                            (cuuint32_t)umma_conv::kBM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(b.swz),
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (b.family == DNNB_KERNEL_UMMA_ROWPAIR && b.rp_linear) {
        // one im2col map per patch plane s: base pixels w in [-3, W], h in {s-1, s-1+TR, ...} (H/TR of them),
        // walked W-fastest, 130 positions per load (conv_rowpair.cuh, LINEAR)
        const int tr = b.rowpair->tr;
        cuuint32_t estr_lin[4] = {1, 1, (cuuint32_t)tr, 1};
        r = CUDA_SUCCESS;
        for (int s = 0; s < tr + 2 && r == CUDA_SUCCESS; ++s) {
            int lower[2] = {-3, s - 1};
            int upper[2] = {1, s - tr};
            r = g_encode_im2col(&b.lin_maps.m[s], dt, 4, const_cast<void *>(in_ptr), dims, strides, lower, upper,
                                (cuuint32_t)b.ic, (cuuint32_t)rowpair::kPatchW, estr_lin, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                swizzle_enum(b.swz), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
    } else if (b.family == DNNB_KERNEL_UMMA_ROWPAIR) {
        cuuint32_t box[4] = {(cuuint32_t)b.ic, (cuuint32_t)rowpair::kPatchW, (cuuint32_t)(b.rowpair->tr + 2), 1};
        r = g_encode_tiled(&b.map_a, dt, 4, const_cast<void *>(in_ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(b.swz), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint32_t box[4] = {(cuuint32_t)bkc, (cuuint32_t)b.up.tw, (cuuint32_t)b.up.th, (cuuint32_t)b.up.nb};
        r = g_encode_tiled(&b.map_a, dt, 4, const_cast<void *>(in_ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(b.swz), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail("tensor-map encode (A, family %d) failed: CUresult %d", b.family, (int)r);
    b.map_a_base = in_ptr;
    return 0;
}

int encode_map_b(Block &b, CUtensorMap *map = nullptr, int bn_rows = 0) {
    if (!map) map = &b.map_b;
    if (!bn_rows) bn_rows = b.bn;
    const int es = b.in_bf16 ? 2 : 4;
    const CUtensorMapDataType dt = b.in_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if (b.family == DNNB_KERNEL_UMMA_ROWPAIR) {
        cuuint64_t dims[3] = {(cuuint64_t)b.ic, (cuuint64_t)(3 * b.oc), 3};
        cuuint64_t strides[2] = {(cuuint64_t)b.ic * es, (cuuint64_t)3 * b.oc * b.ic * es};
        cuuint32_t box[3] = {(cuuint32_t)b.ic, (cuuint32_t)(3 * b.oc), 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = g_encode_tiled(&b.map_b, dt, 3, b.d_wpack, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    swizzle_enum(b.swz), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("tensor-map encode (row-pair weights) failed: CUresult %d", (int)r);
        return 0;
    }
    // [tap][channel chunk][OCp][BKC]: the B tile of a K block (BN rows x one swizzle span) is one
    // contiguous BN*SWZ-byte run, so cold weights stream from HBM instead of being gathered line by line
    const int ocp = b.up.num_n_tiles * b.bn;
    const int bkc = b.swz / es;
    cuuint64_t dims[4] = {(cuuint64_t)bkc, (cuuint64_t)ocp, (cuuint64_t)(b.ic / bkc), (cuuint64_t)(b.kh * b.kw)};
    cuuint64_t strides[3] = {(cuuint64_t)b.swz, (cuuint64_t)ocp * b.swz, (cuuint64_t)(b.ic / bkc) * ocp * b.swz};
    cuuint32_t box[4] = {(cuuint32_t)bkc, (cuuint32_t)bn_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_tiled(map, dt, 4, b.d_wpack, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                swizzle_enum(b.swz), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("tensor-map encode (weights) failed: CUresult %d", (int)r);
    return 0;
}

// scale/shift + LeakyReLU + rounding is monotone non-decreasing when scale > 0: pooling may then
// run on the raw accumulators (f(max(a, b)) == max(f(a), f(b)) exactly)
bool all_scales_positive(const Block &b) {
    for (float v : b.scale)
        if (!(v > 0.f)) return false;
    return true;
}

// Decide how block `i` runs.  in_bf16: storage type of its input.
void plan_block(dnnb_net *net, size_t i, bool in_bf16) {
    Block &b = net->blocks[i];
    const bool last = i + 1 == net->blocks.size();
    const bool want_bf16 = net->precision == DNNB_BF16;
    b.in_bf16 = in_bf16;
    b.out_bf16 = want_bf16 && !last;
    b.round_tf32 = net->precision == DNNB_TF32 && !last;
    b.family = DNNB_KERNEL_SIMT;
    b.pool_fused = false;
    const Knobs &kn = net->knobs;
    const bool force_simt = kn.force_simt;

    // first block straight from the raw 3-channel fp32 image: the patch-resident tcgen05 kernel
    if (i == 0 && !in_bf16 && net->precision != DNNB_FP32 && !force_simt && !kn.no_conv0 && b.ic == 3 &&
        b.oc == 16 && b.kh == 3 && b.kw == 3 && b.same && b.pool_k == 2 && b.pool_s == 2 &&
        b.ih % conv0::kTileRows == 0 && b.iw % 4 == 0) {
        b.family = DNNB_KERNEL_UMMA_CONV0;
        b.pool_fused = true;
        b.pool_separate = false;
        b.op_bf16 = want_bf16;
        conv0::Params &c = b.c0;
        c = conv0::Params{};
        c.B = net->B; c.H = b.ih; c.W = b.iw;
        c.col_tiles = (b.iw + 127) / 128;
        c.num_tiles = net->B * (b.ih / conv0::kTileRows) * c.col_tiles;
        c.leaky = b.leaky; c.round_tf32 = b.round_tf32;
        c.pool_first = all_scales_positive(b) && !kn.no_pool_first;
        b.grid = c.num_tiles < net->num_sms ? c.num_tiles : net->num_sms;
        b.bn = 16; b.swz = 0;
        return;
    }

    // wide few-channel pooled blocks: patch-resident row-pair kernel
    if (net->precision != DNNB_FP32 && !force_simt && !kn.no_rowpair && b.same && b.kh == 3 && b.kw == 3 &&
        b.pool_k == 2 && b.pool_s == 2 && b.ih % 2 == 0 && b.iw % 2 == 0) {
        const int row_bytes = b.ic * (in_bf16 ? 2 : 4);
        const RowpairEntry *e = (row_bytes == 32 || row_bytes == 64 || row_bytes == 128)
                                    ? find_rowpair(in_bf16, b.out_bf16, b.oc, row_bytes, b.ih, kn.rowpair_max_tr) : nullptr;
        if (e) {
            b.family = DNNB_KERNEL_UMMA_ROWPAIR;
            b.rowpair = e;
            b.pool_fused = true;
            b.pool_separate = false;
            b.bn = b.oc; b.swz = row_bytes;
            rowpair::Params &r = b.rp;
            r = rowpair::Params{};
            r.B = net->B; r.H = b.ih; r.W = b.iw;
            r.col_tiles = (b.iw + 127) / 128;
            r.num_tiles = net->B * (b.ih / e->tr) * r.col_tiles;
            // image widths that are a poor multiple of 128 columns (416/208/104: 19 % padding): tile the
            // linearised positions instead (W + 4 per row group, 2-4 % padding)
            const long lin_tiles = ((long)net->B * (b.ih / e->tr) * (b.iw + 4) + 127) / 128;
            b.rp_linear = lin_tiles * 100 < (long)r.num_tiles * 95 && lin_tiles * 128 < (1L << 30);
            if (kn.rowpair_linear >= 0) b.rp_linear = kn.rowpair_linear != 0;
            if (b.rp_linear) { r.num_tiles = (int)lin_tiles; r.col_tiles = 1; }
            r.leaky = b.leaky; r.round_tf32 = b.round_tf32;
            r.pool_first = all_scales_positive(b) && !kn.no_pool_first;
            b.grid = r.num_tiles < net->num_sms ? r.num_tiles : net->num_sms;
            return;
        }
    }

    if (net->precision != DNNB_FP32 && !force_simt && b.same && b.ch == b.ih && b.cw == b.iw) {
        const int es = in_bf16 ? 2 : 4;
        const int row_bytes = b.ic * es;
        const int swz = row_bytes >= 128 ? 128 : row_bytes;
        const bool k_ok = (row_bytes == 32 || row_bytes == 64 || row_bytes % 128 == 0);
        const int bn = b.oc <= 32 ? 32 : b.oc <= 64 ? 64 : b.oc <= 128 ? 128 : 256;
        const int ocp = round_up(b.oc, bn);
        const int c_chunks = row_bytes / swz;
        const int total_kb = b.kh * b.kw * c_chunks;
        const int kbs = (swz < 128 && total_kb % 3 == 0) ? 3 : 1;
        if (k_ok && ocp <= 1024) {
            // fused pool geometry: tile = nb x th x tw, powers of two, product 128, tw <= 16, th >= 2
            bool fuse = b.pool_k == 2 && b.pool_s == 2 && b.ih % 2 == 0 && b.iw % 2 == 0 && b.oc % 8 == 0;
            int tw = 1, th = 1, nb = 1;
            if (fuse) {
                while (tw < 16 && b.iw % (tw * 2) == 0) tw *= 2;
                while (tw * th < 128 && b.ih % (th * 2) == 0) th *= 2;
                nb = 128 / (tw * th);
                fuse = tw >= 2 && th >= 2;
                // a spatial tile spans nb images; with far fewer images in the batch it is mostly padding:
                // take the im2col tiling (consecutive pixels) and a separate pool kernel instead
                if (fuse && net->B * 2 <= nb && nb > kn.lat_fuse_nb && kn.latency) fuse = false;
            }
            UmmaKey key{in_bf16, b.out_bf16, bn, swz, kbs, !fuse, fuse, 1, false};
            const UmmaEntry *e = find_umma(key);
            if (e && !fuse && !b.out_bf16 && ocp == bn && b.oc % 4 != 0 && !kn.no_stage_out) {
                UmmaKey key_so = key;            // narrow odd rows: stage the output tile in smem
                key_so.stage_out = true;
                const UmmaEntry *e_so = find_umma(key_so);
                if (e_so) e = e_so;
            }
            int mt = 1;
            if (e && !kn.no_mt2) {
                // 256-row CTA tiles (two M tiles share each weight tile) when there is enough work for
                // at least two waves of them; DNNB_MT2_MASK (bit i = block i) overrides per block
                const long m_rows = fuse ? (long)((net->B + nb - 1) / nb) * nb * b.ih * b.iw : (long)net->B * b.ih * b.iw;
                const long tiles2 = (m_rows + 255) / 256 * (ocp / bn);
                // measured on B200 (TF32, batch 512): it pays when both accumulator stages survive
                // (BN <= 128: block 3, 0.43 -> 0.28 ms) or when K is long enough to hide the serialised
                // epilogue (block 7, K = 9216: 2.09 -> 1.98 ms; block 6, K = 4608, together with the tail
                // launch below: 1.03 -> 1.00 ms and +1.4 % on the whole power-capped step, because a third
                // less L2->SMEM traffic is also less energy); a loss for blocks 4-5
                const long k_total = (long)b.kh * b.kw * b.ic;
                bool want = tiles2 >= 2L * net->num_sms && (2 * bn <= 256 || k_total >= 4096);
                if (kn.mt2_mask >= 0) want = (kn.mt2_mask >> i) & 1;
                UmmaKey key2 = key;
                key2.mt = 2;
                const UmmaEntry *e2 = want ? find_umma(key2) : nullptr;
                if (e2) { e = e2; mt = 2; }
            }
            if (e) {
                b.umma = e;
                b.family = fuse ? DNNB_KERNEL_UMMA_TILED : DNNB_KERNEL_UMMA_IM2COL;
                b.pool_fused = fuse;
                b.bn = bn;
                b.swz = swz;
                umma_conv::Params &p = b.up;
                p = umma_conv::Params{};
                p.H = b.ih; p.W = b.iw; p.B = net->B; p.OC = b.oc;
                p.KH = b.kh; p.KW = b.kw; p.PT = b.pt; p.PL = b.pl;
                p.c_chunks = c_chunks;
                p.num_n_tiles = ocp / bn;
                if (fuse) {
                    p.tw = tw; p.th = th; p.nb = nb; p.log2_tw = ilog2(tw); p.log2_th = ilog2(th);
                    p.tiles_w = b.iw / tw; p.tiles_h = b.ih / th;
                    p.num_m_tiles = p.tiles_w * p.tiles_h * ((net->B + nb * mt - 1) / (nb * mt));
                    p.OH = b.ih / 2; p.OW = b.iw / 2;
                } else {
                    p.tw = p.th = p.nb = 1;
                    p.M_total = net->B * b.ih * b.iw;
                    p.num_m_tiles = (p.M_total + umma_conv::kBM * mt - 1) / (umma_conv::kBM * mt);
                    p.OH = b.ih; p.OW = b.iw;
                }
                p.leaky = b.leaky;
                p.round_tf32 = b.round_tf32;
                // split-K when the output tiles alone cannot occupy the SMs (small batches)
                const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
                const int kstages = total_kb / kbs;
                p.splits = 1;
                // Splitting costs ~7 us (partials through L2, sibling wait, reduction) and a CTA streams one ring
                // of smem stages per ~2 us round trip, so a K loop of up to three ring passes stays whole
                // (measured, batch 1: block 3, 18 K blocks on 6 stages, 21.5 us split four ways, 14.8 us whole).
                const bool long_k = kstages > 3 * e->stages || kn.split_min_stages > 0;
                if (!fuse && tiles_mn * 2 <= net->num_sms && kstages >= 8 && long_k && kn.latency) {
                    int sp = net->num_sms / tiles_mn;
                    const int min_stages = kn.split_min_stages > 0 ? kn.split_min_stages : 4;
                    if (sp > kstages / min_stages) sp = kstages / min_stages;
                    if (sp >= 2) p.splits = sp;
                }
                const int tiles = tiles_mn * p.splits;       // split-K: one resident CTA per work item (tiles <= SMs)
                b.grid = tiles < net->num_sms ? tiles : net->num_sms;
                // CTA pairs (conv_pair.cuh, tcgen05.mma.cta_group::2) for the big im2col blocks: a third to a
                // quarter less L2->SMEM traffic per flop.  The main launch covers whole waves of pair tiles;
                // the remaining rows go to a tail launch of the single-CTA kernel in 128 x 128 tiles.
                b.pair = nullptr;
                b.umma_tail = nullptr;
                const int sms = net->num_sms;
                if (!fuse && p.splits == 1 && bn == 256 && swz == 128 && kbs == 1 && !e->key.stage_out && ocp == b.oc &&
                    sms % 2 == 0 && !kn.no_pair) {
                    // measured (batch 512, whole step): MT = 1 everywhere 84.4 k img/s, MT = 2 for K >= 4096 83.7 k,
                    // MT = 2 everywhere 82.8 k, no pairs 82.5 k -- two accumulator stages and six smem stages are
                    // worth more than the last quarter of traffic
                    const int pmt = kn.pair_mt;
                    const PairEntry *pe = find_pair(in_bf16, b.out_bf16, pmt);
                    UmmaKey kt{in_bf16, b.out_bf16, 128, swz, kbs, true, false, 1, false};
                    const UmmaEntry *et = find_umma(kt);
                    const int pairs = sms / 2;
                    const int rows_per = 2 * pmt * umma_conv::kBM;
                    const int m_all = p.M_total / rows_per;                 // whole pair tiles
                    const int tiles_p = m_all * p.num_n_tiles;
                    const int min_tiles = kn.pair_min_tiles >= 0 ? kn.pair_min_tiles : 2 * pairs;
                    if (pe && et && tiles_p >= min_tiles && tiles_p > 0) {
                        // stop at the last full wave when the partial one would be less than half full
                        int main_m = m_all;
                        if (tiles_p > pairs && tiles_p % pairs != 0 && (tiles_p % pairs) * 2 < pairs && !kn.pair_no_tail) main_m = (tiles_p / pairs) * pairs / p.num_n_tiles;
                        b.pair = pe;
                        pairconv::Params &q = b.pp;
                        q = pairconv::Params{};
                        q.H = b.ih; q.W = b.iw; q.B = net->B; q.OC = b.oc;
                        q.KH = b.kh; q.KW = b.kw; q.PT = b.pt; q.PL = b.pl;
                        q.c_chunks = c_chunks;
                        q.num_m_tiles = main_m; q.num_n_tiles = p.num_n_tiles;
                        q.M_total = p.M_total; q.m_base = 0;
                        q.leaky = b.leaky; q.round_tf32 = b.round_tf32;
                        const int t = main_m * p.num_n_tiles;
                        b.grid = 2 * (t < pairs ? t : pairs);
                        int rows_done = main_m * rows_per;
                        // whole pair tiles past the full waves: pairs again, in 256 x 128 tiles so that they
                        // spread over twice as many pairs (block 7: 20 tiles -> 40 of the 74 pairs, 100 -> 50 us)
                        b.pair_tail = nullptr;
                        const PairEntry *pt = pmt == 1 ? find_pair(in_bf16, b.out_bf16, 1, 128) : nullptr;
                        const int m_pt = (p.M_total - rows_done) / (2 * umma_conv::kBM);
                        if (pt && m_pt > 0 && !kn.no_pair_tail) {
                            b.pair_tail = pt;
                            b.pp_tail = q;
                            b.pp_tail.m_base = rows_done;
                            b.pp_tail.num_m_tiles = m_pt;
                            b.pp_tail.num_n_tiles = ocp / 128;
                            const int tt = m_pt * (ocp / 128);
                            b.grid_pair_tail = 2 * (tt < pairs ? tt : pairs);
                            rows_done += m_pt * 2 * umma_conv::kBM;
                        }
                        const int rows_left = p.M_total - rows_done;
                        if (rows_left > 0) {
                            const int m_left = (rows_left + umma_conv::kBM - 1) / umma_conv::kBM;
                            b.umma_tail = et;
                            b.up_tail = p;
                            b.up_tail.m_base = rows_done;
                            b.up_tail.num_m_tiles = m_left;
                            b.up_tail.num_n_tiles = ocp / 128;
                            const int tt = m_left * (ocp / 128);
                            b.grid_tail = tt < sms ? tt : sms;
                        }
                    }
                }
                // Without pairs: wave quantisation.  1352 tiles on 148 SMs are 9.14 waves and cost 10.  Cut the
                // block into a main launch of whole waves and a tail launch over the remaining rows with tiles
                // small enough to occupy every SM once.  Every output element still sums its K blocks in the same
                // order, so results do not change (tests: batch invariance, 2 GPUs == 1 GPU).
                // Measured: worth it for the 256 x 256 tiles (block 7: 1.98 -> 1.89 ms); 128 x 256 tiles cut any
                // finer become L2-bandwidth bound and the tail costs as much as the wave it replaces.
                if (!b.pair && !fuse && mt == 2 && p.splits == 1 && !e->key.stage_out && tiles > 2 * sms && tiles % sms != 0 &&
                    !kn.no_tail) {
                    const int main_m = (tiles / sms) * sms / p.num_n_tiles;           // whole m-tiles in the full waves
                    const int rows_left = p.M_total - main_m * mt * umma_conv::kBM;
                    const int m_left = (rows_left + umma_conv::kBM - 1) / umma_conv::kBM;
                    UmmaKey kt{in_bf16, b.out_bf16, 128, swz, kbs, true, false, 1, false};
                    const UmmaEntry *et = find_umma(kt);
                    const int t_tiles = m_left * (ocp / 128);
                    if (et && 128 < bn * mt && ocp % 128 == 0 && t_tiles <= sms) {
                        b.umma_tail = et;
                        b.up_tail = p;
                        b.up_tail.m_base = main_m * mt * umma_conv::kBM;
                        b.up_tail.num_m_tiles = m_left;
                        b.up_tail.num_n_tiles = ocp / 128;
                        b.grid_tail = t_tiles;
                        p.num_m_tiles = main_m;
                    }
                }
            }
        }
    }
    if (b.family == DNNB_KERNEL_SIMT)
        b.pool_fused = b.pool_k == 2 && b.pool_s == 2 && b.ch % 2 == 0 && b.cw % 2 == 0;
    b.pool_separate = b.pool_k > 0 && !b.pool_fused;
}

template <bool BF16, bool OUT_BF16, bool IN_U8 = false>
cudaError_t launch_conv0_t(const conv0::Params &p, int grid, cudaStream_t st) {
    auto kern = conv0::conv0_kernel<BF16, OUT_BF16, IN_U8>;
    const int smem = conv0::kSmemBytes + 128;
    static bool done[64] = {};
    cudaError_t e = set_smem_once(kern, smem, done);
    if (e != cudaSuccess) return e;
    return launch_ex(kern, (unsigned)grid, conv0::kThreads, smem, st, false, p);
}
cudaError_t launch_conv0(const conv0::Params &p, bool bf16, bool out_bf16, int grid, cudaStream_t st, bool in_u8 = false) {
    if (in_u8) {
        if (bf16) return out_bf16 ? launch_conv0_t<true, true, true>(p, grid, st) : launch_conv0_t<true, false, true>(p, grid, st);
        return out_bf16 ? launch_conv0_t<false, true, true>(p, grid, st) : launch_conv0_t<false, false, true>(p, grid, st);
    }
    if (bf16) return out_bf16 ? launch_conv0_t<true, true>(p, grid, st) : launch_conv0_t<true, false>(p, grid, st);
    return out_bf16 ? launch_conv0_t<false, true>(p, grid, st) : launch_conv0_t<false, false>(p, grid, st);
}

template <typename TIn, typename TOut>
cudaError_t launch_simt_conv(const simt::ConvParams &p, cudaStream_t st) {
    dim3 grid((unsigned)(p.B * p.tiles_y * p.tiles_x), (unsigned)((p.OC + 63) / 64));
    simt::conv_fused_kernel<TIn, TOut><<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t run_simt_conv(const simt::ConvParams &p, bool in_bf16, bool out_bf16, cudaStream_t st) {
    if (in_bf16) return out_bf16 ? launch_simt_conv<__nv_bfloat16, __nv_bfloat16>(p, st) : launch_simt_conv<__nv_bfloat16, float>(p, st);
    return out_bf16 ? launch_simt_conv<float, __nv_bfloat16>(p, st) : launch_simt_conv<float, float>(p, st);
}

int ew_grid(size_t n, int num_sms) {
    size_t blocks = (n + 255) / 256;
    size_t cap = (size_t)num_sms * 16;
    return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

cudaError_t run_pool(const simt::PoolParams &p, bool bf16, int num_sms, cudaStream_t st) {
    const size_t n = (size_t)p.B * p.OH * p.OW * p.C;
    if (p.C % 4 == 0 && p.KH == 2 && p.KW == 2 && p.SH == 1 && p.SW == 1 && p.PT == 0 && p.PL == 0 && p.OH == p.IH &&
        p.OW == p.IW && (long)p.B * p.OH < 2000000000L) {
        const unsigned threads = p.C / 4 < 256 ? (p.C / 4 + 31) / 32 * 32 : 256;
        if (bf16) return launch_ex(simt::maxpool_2x2s1_kernel<__nv_bfloat16>, (unsigned)(p.B * p.OH), threads, 0, st, false, p);
        return launch_ex(simt::maxpool_2x2s1_kernel<float>, (unsigned)(p.B * p.OH), threads, 0, st, false, p);
    }
    if (p.C % 4 == 0) {
        const unsigned grid = ew_grid(n / 4, num_sms);
        if (bf16) return launch_ex(simt::maxpool_vec4_kernel<__nv_bfloat16>, grid, 256u, 0, st, false, p);
        return launch_ex(simt::maxpool_vec4_kernel<float>, grid, 256u, 0, st, false, p);
    }
    const unsigned grid = ew_grid(n, num_sms);
    if (bf16) return launch_ex(simt::maxpool_kernel<__nv_bfloat16>, grid, 256u, 0, st, false, p);
    return launch_ex(simt::maxpool_kernel<float>, grid, 256u, 0, st, false, p);
}

// launches the kernels of one block; `in` has the block's input storage type
int run_block(dnnb_net *net, Block &b, const void *in, void *final_out_override, cudaStream_t st,
              float *prepool = nullptr) {
    void *conv_out = b.d_conv_out;
    void *out = final_out_override ? final_out_override : b.d_out;
    if (!b.pool_separate) conv_out = out;
    if (b.family == DNNB_KERNEL_SIMT) {
        simt::ConvParams p{};
        p.in = in; p.w = b.d_kernel; p.out = conv_out; p.scale = b.d_scale; p.shift = b.d_shift;
        p.B = net->B; p.IH = b.ih; p.IW = b.iw; p.IC = b.ic; p.OH = b.ch; p.OW = b.cw; p.OC = b.oc;
        p.KH = b.kh; p.KW = b.kw; p.SH = 1; p.SW = 1; p.PT = b.pt; p.PL = b.pl;
        p.leaky = b.leaky; p.pool = b.pool_fused; p.round_tf32 = b.round_tf32;
        p.tiles_y = (b.ch + 7) / 8; p.tiles_x = (b.cw + 7) / 8;
        CK(run_simt_conv(p, b.in_bf16, b.out_bf16, st));
    } else if (b.family == DNNB_KERNEL_UMMA_CONV0) {
        conv0::Params p = b.c0;
        p.in = in; p.out = conv_out; p.wpack = b.d_wpack;
        p.scale = b.d_scale; p.shift = b.d_shift; p.prepool = prepool;
        const bool u8 = net->cur_in_u8 && &b == &net->blocks.front();
        if (u8) { p.lut = net->d_lut_conv0; p.swap_rb = net->cur_swap_rb; }
        CK(launch_conv0(p, b.op_bf16, b.out_bf16, b.grid, st, u8));
    } else if (b.family == DNNB_KERNEL_UMMA_ROWPAIR) {
        if (b.map_a_base != in) {
            if (encode_map_a(net, b, in)) return 1;
        }
        rowpair::Params p = b.rp;
        p.scale = b.d_scale; p.shift = b.d_shift; p.out = conv_out; p.prepool = prepool;
        CK((b.rp_linear ? b.rowpair->launch_linear : b.rowpair->launch)(b.map_a, b.map_b, p, b.lin_maps, b.grid, st));
    } else {
        if (b.map_a_base != in) {
            if (encode_map_a(net, b, in)) return 1;
        }
        umma_conv::Params p = b.up;
        p.scale = b.d_scale; p.shift = b.d_shift; p.out = conv_out; p.prepool = prepool;
        p.workspace = b.d_workspace; p.counters = b.d_counters;
        if (b.pair) {
            pairconv::Params q = b.pp;
            q.scale = b.d_scale; q.shift = b.d_shift; q.out = conv_out;
            if (q.num_m_tiles > 0) CK(b.pair->launch(b.map_a, b.map_b_pair, q, b.grid, st));
            if (b.pair_tail) {
                pairconv::Params t = b.pp_tail;
                t.scale = b.d_scale; t.shift = b.d_shift; t.out = conv_out;
                CK(b.pair_tail->launch(b.map_a, b.map_b_pair_tail, t, b.grid_pair_tail, st));
            }
        } else {
            CK(b.umma->launch(b.map_a, b.map_b, p, p.splits > 1 ? -b.grid : b.grid, st));
        }
        if (b.umma_tail) {
            umma_conv::Params q = b.up_tail;
            q.scale = b.d_scale; q.shift = b.d_shift; q.out = conv_out; q.prepool = prepool;
            CK(b.umma_tail->launch(b.map_a, b.map_b_tail, q, b.grid_tail, st));
        }
    }
    if (b.pool_separate) {
        simt::PoolParams q{};
        q.in = conv_out; q.out = out; q.B = net->B; q.IH = b.ch; q.IW = b.cw; q.C = b.oc; q.OH = b.oh; q.OW = b.ow;
        q.KH = q.KW = b.pool_k; q.SH = q.SW = b.pool_s; q.PT = b.ppt; q.PL = b.ppl;
        CK(run_pool(q, b.out_bf16, net->num_sms, st));
    }
    return 0;
}

// in_u8: dev_in holds uint8 frames of the network's size and block 0 is the conv0 kernel, which then
// normalises them in its own loader (no separate preprocessing kernel, 3 instead of 12 bytes per pixel read)
struct PdlScope {
    bool prev;
    explicit PdlScope(bool on) : prev(t_pdl) { t_pdl = on; }
    ~PdlScope() { t_pdl = prev; }
};

int forward_impl(dnnb_net *net, const void *dev_in, void *dev_out, cudaStream_t st, bool in_u8 = false, int swap_rb = 0) {
    PdlScope pdl(net->pdl);
    const void *cur = dev_in;
    Block &first = net->blocks.front();
    net->cur_in_u8 = in_u8;
    net->cur_swap_rb = swap_rb;
    if (first.family != DNNB_KERNEL_SIMT && first.family != DNNB_KERNEL_UMMA_CONV0) {
        // a tcgen05 first block reads its storage type: TF32-rounded fp32, or bf16
        const size_t n = (size_t)net->B * net->H * net->W * net->C;
        const int grid = ew_grid(n, net->num_sms);
        if (first.in_bf16)
            simt::cast_input_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const float *)dev_in, (__nv_bfloat16 *)net->d_in_cast, n, 0);
        else
            simt::cast_input_kernel<float><<<grid, 256, 0, st>>>((const float *)dev_in, (float *)net->d_in_cast, n, 1);
        CK(cudaGetLastError());
        cur = net->d_in_cast;
    }
    const size_t n_run = net->block_limit > 0 && (size_t)net->block_limit < net->blocks.size() ? (size_t)net->block_limit
                                                                                               : net->blocks.size();
    for (size_t i = 0; i < n_run; ++i) {
        Block &b = net->blocks[i];
        const bool last = i + 1 == net->blocks.size();
        void *override_out = (last && dev_out && !b.out_bf16) ? dev_out : nullptr;
        const bool prof = net->prof_block == (int)i && net->prof_used < 8192;
        if (prof) {
            if (net->prof_used == net->prof_events.size()) {
                cudaEvent_t a, c;
                CK(cudaEventCreate(&a));
                CK(cudaEventCreate(&c));
                net->prof_events.emplace_back(a, c);
            }
            CK(cudaEventRecord(net->prof_events[net->prof_used].first, st));
        }
        if (run_block(net, b, cur, override_out, st)) return 1;
        if (prof) {
            CK(cudaEventRecord(net->prof_events[net->prof_used].second, st));
            net->prof_used++;
        }
        cur = override_out ? override_out : b.d_out;
    }
    return 0;
}

// Records one forward on `st` into a graph (stream capture; nothing executes).  Returns non-zero, with the
// stream back in its normal state, when the capture cannot be made.
int capture_forward(dnnb_net *net, const void *dev_in, void *dev_out, cudaStream_t st, cudaGraphExec_t *exec) {
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    const int rc = forward_impl(net, dev_in, dev_out, st);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (rc != 0 || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return 1;
    }
    const cudaError_t ei = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    return 0;
}

int check_net(dnnb_net *net, bool need_final) {
    if (!net) return fail("null network handle");
    if (need_final && !net->finalized) return fail("network not finalized");
    return 0;
}

}  // namespace

// =============================================================================================
extern "C" {

const char *dnnb_last_error(void) { return g_err.c_str(); }
int dnnb_version(void) { return 100; }

int dnnb_device_count(int *count) {
    if (!count) return fail("null argument");
    CK(cudaGetDeviceCount(count));
    return 0;
}

// ---------------------------------------------------------------------------- op-level
namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        CK(cudaMalloc(&p, bytes ? bytes : 4));
        return 0;
    }
    float *f() { return static_cast<float *>(p); }
};
int sm_count() {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}
}  // namespace

int conv2d_mul(float *in_layer, float *col, float *kernel_r, float *result, int batch, int oh, int ow, int od,
               int ih, int iw, int ic, int kh, int kw, int sh, int sw) {
    (void)col;
    if (!in_layer || !kernel_r || !result) return fail("conv2d_mul: null pointer");
    const size_t n_in = (size_t)batch * ih * iw * ic, n_out = (size_t)batch * oh * ow * od;
    const size_t n_k = (size_t)kh * kw * ic * od;
    // kernel_r[(c*kh*kw + di*kw + dj)][d]  ->  HWIO
    std::vector<float> hwio(n_k);
    for (int c = 0; c < ic; ++c)
        for (int t = 0; t < kh * kw; ++t)
            memcpy(&hwio[((size_t)t * ic + c) * od], &kernel_r[((size_t)c * kh * kw + t) * od], sizeof(float) * od);
    DevBuf din, dk, dout;
    if (din.alloc(n_in * 4) || dk.alloc(n_k * 4) || dout.alloc(n_out * 4)) return 1;
    CK(cudaMemcpy(din.p, in_layer, n_in * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dk.p, hwio.data(), n_k * 4, cudaMemcpyHostToDevice));
    simt::ConvParams p{};
    p.in = din.p; p.w = dk.f(); p.out = dout.p;
    p.B = batch; p.IH = ih; p.IW = iw; p.IC = ic; p.OH = oh; p.OW = ow; p.OC = od;
    p.KH = kh; p.KW = kw; p.SH = sh; p.SW = sw;
    p.tiles_y = (oh + 7) / 8; p.tiles_x = (ow + 7) / 8;
    CK(launch_simt_conv<float, float>(p, nullptr));
    CK(cudaMemcpy(result, dout.p, n_out * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int bias_add(float *in_layer, float *biases, float *result, int shape_b, int shape_h, int shape_w, int shape_d) {
    if (!in_layer || !biases || !result) return fail("bias_add: null pointer");
    const size_t n = (size_t)shape_b * shape_h * shape_w * shape_d;
    DevBuf din, db, dout;
    if (din.alloc(n * 4) || db.alloc((size_t)shape_d * 4) || dout.alloc(n * 4)) return 1;
    CK(cudaMemcpy(din.p, in_layer, n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db.p, biases, (size_t)shape_d * 4, cudaMemcpyHostToDevice));
    simt::bias_add_kernel<<<ew_grid(n, sm_count()), 256>>>(din.f(), db.f(), dout.f(), n, shape_d);
    CK(cudaGetLastError());
    CK(cudaMemcpy(result, dout.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int batch_norm(float *in_layer, float *mean, float *variance, float *gamma, float epsilon, float *result, int batch,
               int oh, int ow, int od) {
    if (!in_layer || !mean || !variance || !gamma || !result) return fail("batch_norm: null pointer");
    const size_t n = (size_t)batch * oh * ow * od;
    std::vector<float> sd(od);
    for (int d = 0; d < od; ++d) sd[d] = (float)sqrt((double)(variance[d] + epsilon));   // dnn_openblas.c:48-50
    DevBuf din, dm, ds, dg, dout;
    if (din.alloc(n * 4) || dm.alloc((size_t)od * 4) || ds.alloc((size_t)od * 4) || dg.alloc((size_t)od * 4) ||
        dout.alloc(n * 4))
        return 1;
    CK(cudaMemcpy(din.p, in_layer, n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dm.p, mean, (size_t)od * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ds.p, sd.data(), (size_t)od * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dg.p, gamma, (size_t)od * 4, cudaMemcpyHostToDevice));
    simt::batch_norm_kernel<<<ew_grid(n, sm_count()), 256>>>(din.f(), dm.f(), ds.f(), dg.f(), dout.f(), n, od);
    CK(cudaGetLastError());
    CK(cudaMemcpy(result, dout.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int max_pool2d(float *in_layer, float *result, int batch, int oh, int ow, int od, int ih, int iw, int ic, int kh,
               int kw, int sh, int sw) {
    if (!in_layer || !result) return fail("max_pool2d: null pointer");
    if (od != ic) return fail("max_pool2d: od != ic");
    if ((oh - 1) * sh + kh > ih || (ow - 1) * sw + kw > iw) return fail("max_pool2d: window leaves the padded input");
    const size_t n_in = (size_t)batch * ih * iw * ic, n_out = (size_t)batch * oh * ow * od;
    DevBuf din, dout;
    if (din.alloc(n_in * 4) || dout.alloc(n_out * 4)) return 1;
    CK(cudaMemcpy(din.p, in_layer, n_in * 4, cudaMemcpyHostToDevice));
    simt::PoolParams q{};
    q.in = din.p; q.out = dout.p; q.B = batch; q.IH = ih; q.IW = iw; q.C = ic; q.OH = oh; q.OW = ow;
    q.KH = kh; q.KW = kw; q.SH = sh; q.SW = sw;
    CK(run_pool(q, false, sm_count(), nullptr));
    CK(cudaMemcpy(result, dout.p, n_out * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int leaky_relu(float *in_layer, float *result, int batch, int oh, int ow, int od) {
    if (!in_layer || !result) return fail("leaky_relu: null pointer");
    const size_t n = (size_t)batch * oh * ow * od;
    DevBuf din, dout;
    if (din.alloc(n * 4) || dout.alloc(n * 4)) return 1;
    CK(cudaMemcpy(din.p, in_layer, n * 4, cudaMemcpyHostToDevice));
    simt::leaky_relu_kernel<<<ew_grid(n, sm_count()), 256>>>(din.f(), dout.f(), n);
    CK(cudaGetLastError());
    CK(cudaMemcpy(result, dout.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// ---------------------------------------------------------------------------- network
int dnnb_net_create(int device, int precision, int batch, int in_h, int in_w, int in_c, dnnb_net **out) {
    if (!out) return fail("null out pointer");
    if (precision != DNNB_TF32 && precision != DNNB_BF16 && precision != DNNB_FP32) return fail("bad precision %d", precision);
    if (batch <= 0 || in_h <= 0 || in_w <= 0 || in_c <= 0) return fail("bad input shape");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail("device %d not present (%d devices)", device, ndev);
    DeviceGuard g(device);
    int major = 0, sms = 0;
    CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (major != 10) return fail("device %d is sm_%d0: this library is built for sm_100a only", device, major);
    dnnb_net *net = new dnnb_net();
    net->device = device; net->precision = precision;
    net->B = batch; net->H = in_h; net->W = in_w; net->C = in_c;
    net->num_sms = sms;
    cudaError_t e = cudaStreamCreateWithFlags(&net->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete net; return fail("cudaStreamCreate: %s", cudaGetErrorString(e)); }
    *out = net;
    return 0;
}

int dnnb_net_add_block(dnnb_net *net, const float *kernel, int kh, int kw, int ic, int oc, int padding_same,
                       const float *biases, const float *mean, const float *variance, const float *gamma,
                       float epsilon, int leaky, int pool_k, int pool_s, int pool_same) {
    if (check_net(net, false)) return 1;
    if (net->finalized) return fail("network already finalized");
    if (!kernel || kh <= 0 || kw <= 0 || ic <= 0 || oc <= 0) return fail("bad kernel");
    const bool bn = mean || variance || gamma;
    if (bn && !(mean && variance && gamma)) return fail("batch norm needs mean, variance and gamma");
    if (pool_k < 0 || (pool_k > 0 && pool_s <= 0)) return fail("bad pool");
    int in_c = net->blocks.empty() ? net->C : net->blocks.back().oc;
    if (ic != in_c) return fail("block %zu: kernel has %d input channels, input has %d", net->blocks.size(), ic, in_c);
    Block b;
    b.kh = kh; b.kw = kw; b.ic = ic; b.oc = oc; b.same = padding_same ? 1 : 0;
    b.has_bn = bn; b.leaky = leaky != 0;
    b.pool_k = pool_k; b.pool_s = pool_s; b.pool_same = pool_same ? 1 : 0;
    b.kernel.assign(kernel, kernel + (size_t)kh * kw * ic * oc);
    b.scale.resize(oc); b.shift.resize(oc);
    for (int d = 0; d < oc; ++d) {
        // fold  ((x + bias - mean) / sqrt(var + eps)) * gamma   into   x * scale + shift
        double sc = 1.0, sh = biases ? (double)biases[d] : 0.0;
        if (bn) {
            const float sd = (float)sqrt((double)(variance[d] + epsilon));   // dnn_openblas.c:48-50
            sc = (double)gamma[d] / (double)sd;
            sh = (sh - (double)mean[d]) * sc;
        }
        b.scale[d] = (float)sc;
        b.shift[d] = (float)sh;
    }
    net->blocks.push_back(std::move(b));
    return 0;
}

int dnnb_net_finalize(dnnb_net *net) {
    if (check_net(net, false)) return 1;
    if (net->finalized) return fail("network already finalized");
    if (net->blocks.empty()) return fail("network has no blocks");
    DeviceGuard g(net->device);
    if (load_driver_fns()) return 1;
    int h = net->H, w = net->W;
    bool in_bf16 = false;
    net->launches = 0;
    net->knobs = read_knobs();
    net->pdl = net->knobs.latency && !net->knobs.no_pdl;
    net->use_graph = net->knobs.latency && !net->knobs.no_graph;
    for (size_t i = 0; i < net->blocks.size(); ++i) {
        Block &b = net->blocks[i];
        b.ih = h; b.iw = w;
        int pb;
        out_pads(h, b.kh, 1, b.same, &b.ch, &b.pt, &pb);
        out_pads(w, b.kw, 1, b.same, &b.cw, &b.pl, &pb);
        if (b.ch <= 0 || b.cw <= 0) return fail("block %zu: empty conv output", i);
        b.oh = b.ch; b.ow = b.cw;
        if (b.pool_k > 0) {
            out_pads(b.ch, b.pool_k, b.pool_s, b.pool_same, &b.oh, &b.ppt, &pb);
            out_pads(b.cw, b.pool_k, b.pool_s, b.pool_same, &b.ow, &b.ppl, &pb);
            if (b.oh <= 0 || b.ow <= 0) return fail("block %zu: empty pool output", i);
        }
        if (i == 0 && net->precision == DNNB_BF16) {
            // conv0 reads the fp32 image itself; any other tcgen05 first block gets a bf16 cast
            plan_block(net, i, false);
            if (b.family != DNNB_KERNEL_UMMA_CONV0) {
                plan_block(net, i, true);
                if (b.family == DNNB_KERNEL_SIMT) plan_block(net, i, false);
            }
        } else {
            plan_block(net, i, in_bf16);
        }
        const size_t es_out = b.out_bf16 ? 2 : 4;
        b.out_elems = (size_t)net->B * b.oh * b.ow * b.oc;
        b.conv_out_elems = b.pool_fused ? b.out_elems : (size_t)net->B * b.ch * b.cw * b.oc;
        CK(cudaMalloc(&b.d_out, b.out_elems * es_out));
        if (b.pool_separate) CK(cudaMalloc(&b.d_conv_out, b.conv_out_elems * es_out));
        else b.d_conv_out = b.d_out;
        // epilogue constants, padded to the N tiling
        const int ocp = (b.family == DNNB_KERNEL_UMMA_TILED || b.family == DNNB_KERNEL_UMMA_IM2COL) ? b.up.num_n_tiles * b.bn : b.oc;
        std::vector<float> sc(ocp, 0.f), sh(ocp, 0.f);
        memcpy(sc.data(), b.scale.data(), sizeof(float) * b.oc);
        memcpy(sh.data(), b.shift.data(), sizeof(float) * b.oc);
        CK(cudaMalloc(&b.d_scale, sizeof(float) * ocp));
        CK(cudaMalloc(&b.d_shift, sizeof(float) * ocp));
        CK(cudaMemcpy(b.d_scale, sc.data(), sizeof(float) * ocp, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(b.d_shift, sh.data(), sizeof(float) * ocp, cudaMemcpyHostToDevice));
        if (b.family == DNNB_KERNEL_UMMA_CONV0) {
            std::vector<uint8_t> pk(conv0::kWBytes);
            if (b.op_bf16) conv0::pack_weights<uint16_t>(b.kernel.data(), reinterpret_cast<uint16_t *>(pk.data()), host_bf16);
            else conv0::pack_weights<float>(b.kernel.data(), reinterpret_cast<float *>(pk.data()), host_round_tf32);
            CK(cudaMalloc(&b.d_wpack, conv0::kWBytes));
            CK(cudaMemcpy(b.d_wpack, pk.data(), conv0::kWBytes, cudaMemcpyHostToDevice));
        } else if (b.family == DNNB_KERNEL_UMMA_ROWPAIR) {
            // [dj][(2-di)*OC + oc][IC]: for every dj the rows are W(di=2) | W(di=1) | W(di=0)
            const size_t n = (size_t)9 * b.oc * b.ic;
            std::vector<float> pk(n);
            for (int di = 0; di < 3; ++di)
                for (int dj = 0; dj < 3; ++dj)
                    for (int c = 0; c < b.ic; ++c)
                        for (int d = 0; d < b.oc; ++d)
                            pk[((size_t)dj * 3 * b.oc + (size_t)(2 - di) * b.oc + d) * b.ic + c] =
                                b.kernel[((size_t)(di * 3 + dj) * b.ic + c) * b.oc + d];
            if (b.in_bf16) {
                std::vector<uint16_t> h(n);
                for (size_t q = 0; q < n; ++q) h[q] = host_bf16(pk[q]);
                CK(cudaMalloc(&b.d_wpack, n * 2));
                CK(cudaMemcpy(b.d_wpack, h.data(), n * 2, cudaMemcpyHostToDevice));
            } else {
                for (size_t q = 0; q < n; ++q) pk[q] = host_round_tf32(pk[q]);
                CK(cudaMalloc(&b.d_wpack, n * 4));
                CK(cudaMemcpy(b.d_wpack, pk.data(), n * 4, cudaMemcpyHostToDevice));
            }
            if (encode_map_b(b)) return 1;
        } else if (b.family == DNNB_KERNEL_SIMT) {
            CK(cudaMalloc(&b.d_kernel, sizeof(float) * b.kernel.size()));
            CK(cudaMemcpy(b.d_kernel, b.kernel.data(), sizeof(float) * b.kernel.size(), cudaMemcpyHostToDevice));
        } else {
            if (b.family != DNNB_KERNEL_UMMA_TILED && b.family != DNNB_KERNEL_UMMA_IM2COL) return fail("internal: unknown kernel family");
            // prepack HWIO -> [tap][channel chunk][OCp][BKC], K-major, rounded once to the MMA operand type
            const int taps = b.kh * b.kw;
            const size_t n = (size_t)taps * ocp * b.ic;
            const int bkc = b.swz / (b.in_bf16 ? 2 : 4), chunks = b.ic / bkc;
            auto widx = [&](int t, int c, int d) {
                return (((size_t)t * chunks + c / bkc) * ocp + d) * bkc + c % bkc;
            };
            if (b.in_bf16) {
                std::vector<uint16_t> pk(n, 0);
                for (int t = 0; t < taps; ++t)
                    for (int c = 0; c < b.ic; ++c)
                        for (int d = 0; d < b.oc; ++d)
                            pk[widx(t, c, d)] = host_bf16(b.kernel[((size_t)t * b.ic + c) * b.oc + d]);
                CK(cudaMalloc(&b.d_wpack, n * 2));
                CK(cudaMemcpy(b.d_wpack, pk.data(), n * 2, cudaMemcpyHostToDevice));
            } else {
                std::vector<float> pk(n, 0.f);
                for (int t = 0; t < taps; ++t)
                    for (int c = 0; c < b.ic; ++c)
                        for (int d = 0; d < b.oc; ++d)
                            pk[widx(t, c, d)] = host_round_tf32(b.kernel[((size_t)t * b.ic + c) * b.oc + d]);
                CK(cudaMalloc(&b.d_wpack, n * 4));
                CK(cudaMemcpy(b.d_wpack, pk.data(), n * 4, cudaMemcpyHostToDevice));
            }
            if (encode_map_b(b)) return 1;
            if (b.umma_tail) {
                if (encode_map_b(b, &b.map_b_tail, b.umma_tail->key.bn)) return 1;
                net->launches += 1;
            }
            if (b.pair && encode_map_b(b, &b.map_b_pair, 128)) return 1;
            if (b.pair_tail) {
                if (encode_map_b(b, &b.map_b_pair_tail, 64)) return 1;
                net->launches += 1;
            }
            if (b.up.splits > 1) {
                const size_t ws = (size_t)b.up.splits * b.up.num_m_tiles * (b.umma->key.mt) * umma_conv::kBM * ocp * sizeof(float);
                const size_t nc = (size_t)b.up.num_m_tiles * b.up.num_n_tiles * 2 * sizeof(int);
                CK(cudaMalloc(&b.d_workspace, ws));
                CK(cudaMalloc(&b.d_counters, nc));
                CK(cudaMemset(b.d_workspace, 0, ws));
                CK(cudaMemset(b.d_counters, 0, nc));
            }
        }
        net->launches += 1 + (b.pool_separate ? 1 : 0);
        in_bf16 = b.out_bf16;
        h = b.oh; w = b.ow;
    }
    const size_t n_in = (size_t)net->B * net->H * net->W * net->C;
    CK(cudaMalloc(&net->d_in2[0], n_in * 4));
    CK(cudaMalloc(&net->d_in2[1], n_in * 4));
    net->d_in = net->d_in2[0];
    CK(cudaStreamCreateWithFlags(&net->copy_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&net->copy_out, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        CK(cudaEventCreateWithFlags(&net->in_ready[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&net->in_free[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&net->out_ready[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&net->out_free[k], cudaEventDisableTiming));
    }
    Block &first = net->blocks.front();
    if (first.family != DNNB_KERNEL_SIMT && first.family != DNNB_KERNEL_UMMA_CONV0) {
        CK(cudaMalloc(&net->d_in_cast, n_in * (first.in_bf16 ? 2 : 4)));
        net->launches += 1;
        if (encode_map_a(net, first, net->d_in_cast)) return 1;
    }
    for (size_t i = 1; i < net->blocks.size(); ++i)
        if (net->blocks[i].family == DNNB_KERNEL_UMMA_TILED || net->blocks[i].family == DNNB_KERNEL_UMMA_IM2COL ||
            net->blocks[i].family == DNNB_KERNEL_UMMA_ROWPAIR)
            if (encode_map_a(net, net->blocks[i], net->blocks[i - 1].d_out)) return 1;
    Block &lastb = net->blocks.back();
    if (lastb.out_bf16) return fail("internal: last block stored as bf16");
    net->finalized = true;
    return 0;
}

int dnnb_net_num_blocks(dnnb_net *net, int *n) {
    if (check_net(net, false) || !n) return fail("bad argument");
    *n = (int)net->blocks.size();
    return 0;
}

int dnnb_net_out_shape(dnnb_net *net, int *b, int *h, int *w, int *c) {
    if (check_net(net, true)) return 1;
    const Block &l = net->blocks.back();
    if (b) *b = net->B;
    if (h) *h = l.oh;
    if (w) *w = l.ow;
    if (c) *c = l.oc;
    return 0;
}

int dnnb_net_forward_device(dnnb_net *net, const void *dev_in, void *dev_out, void *stream) {
    if (check_net(net, true)) return 1;
    if (!dev_in) return fail("null device input");
    DeviceGuard g(net->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : net->stream;
    net->last_final = dev_out ? nullptr : static_cast<const float *>(net->blocks.back().d_out);
    net->last_in = dev_in;
    if (net->use_graph && net->prof_block < 0) {
        for (dnnb_net::GraphSlot &g : net->graphs)
            if (g.in == dev_in && g.out == dev_out) {
                g.used = ++net->graph_tick;
                CK(cudaGraphLaunch(g.exec, st));
                return 0;
            }
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (net->warm_in == dev_in && net->warm_out == dev_out && cudaStreamIsCapturing(st, &cs) == cudaSuccess &&
            cs == cudaStreamCaptureStatusNone) {
            cudaGraphExec_t exec = nullptr;
            if (capture_forward(net, dev_in, dev_out, st, &exec) == 0) {
                if (net->graphs.size() >= 4) {            // keep the four most recently used pointer pairs
                    size_t victim = 0;
                    for (size_t i = 1; i < net->graphs.size(); ++i)
                        if (net->graphs[i].used < net->graphs[victim].used) victim = i;
                    cudaGraphExecDestroy(net->graphs[victim].exec);
                    net->graphs.erase(net->graphs.begin() + victim);
                }
                dnnb_net::GraphSlot g;
                g.in = dev_in; g.out = dev_out; g.exec = exec; g.used = ++net->graph_tick;
                net->graphs.push_back(g);
                CK(cudaGraphLaunch(exec, st));
                return 0;
            }
            net->use_graph = false;       // capture is not available here (legacy stream, driver): launch directly from now on
        }
        net->warm_in = dev_in;
        net->warm_out = dev_out;
    }
    return forward_impl(net, dev_in, dev_out, st);
}

namespace {
cudaError_t dnnb_sync_streams(dnnb_net *net) {
    cudaError_t e = cudaStreamSynchronize(net->stream);
    if (e == cudaSuccess && net->copy_in) e = cudaStreamSynchronize(net->copy_in);
    if (e == cudaSuccess && net->copy_out) e = cudaStreamSynchronize(net->copy_out);
    return e;
}

// One axis of cv::resize's INTER_LINEAR set-up for 8-bit images (OpenCV resize.cpp), in its
// arithmetic: double scale, float32 fractional position, weights short(round_half_even(w * 2048)).
// `clamp_x`: the x-axis border rule (offset clamped, weight reset); rows are clipped at fetch time.
void resize_axis_table(int src, int dst, bool clamp_x, int *ofs, int *wts) {
    const double scale = 1.0 / ((double)dst / (double)src);
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (clamp_x) {
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= src - 1) { s = src - 1; f = 0.f; }
        }
        ofs[d] = s;
        const int w0 = (int)lrintf((1.f - f) * 2048.f), w1 = (int)lrintf(f * 2048.f);
        wts[d] = (w0 & 0xffff) | (w1 << 16);
    }
}

int ensure_u8(dnnb_net *net, int src_h = 0, int src_w = 0) {
    if (net->C != 3) return fail("uint8 frames need a 3-channel network input");
    if (src_h <= 0) src_h = net->H;
    if (src_w <= 0) src_w = net->W;
    const size_t n = (size_t)net->B * src_h * src_w * 3;
    if (src_h == net->H && src_w == net->W && n % 12 != 0) return fail("uint8 frames: B*H*W must be a multiple of 4");
    if (!net->d_lut) {
        float lut[256];
        for (int i = 0; i < 256; ++i) lut[i] = (float)((double)i / 255.0);   // numpy: uint8 / 255. -> float64 -> float32
        CK(cudaMalloc(&net->d_lut, sizeof(lut)));
        CK(cudaMemcpy(net->d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
        // block 0's operand bits for the same values: what cast + rounding in the fp32 path produce
        uint32_t bits[256];
        const bool bf16 = net->blocks.front().op_bf16;
        for (int i = 0; i < 256; ++i) {
            uint32_t u;
            memcpy(&u, &lut[i], 4);
            bits[i] = bf16 ? (uint32_t)host_bf16(lut[i]) : u + 0x1000u;
        }
        CK(cudaMalloc(&net->d_lut_conv0, sizeof(bits)));
        CK(cudaMemcpy(net->d_lut_conv0, bits, sizeof(bits), cudaMemcpyHostToDevice));
    }
    if (n > net->u8_capacity) {
        CK(dnnb_sync_streams(net));          // earlier forwards may still read the old staging buffers
        for (int k = 0; k < 2; ++k) {
            if (net->d_u8[k]) CK(cudaFree(net->d_u8[k]));
            net->d_u8[k] = nullptr;
            CK(cudaMalloc(&net->d_u8[k], n));
        }
        net->u8_capacity = n;
    }
    if ((src_h != net->H || src_w != net->W) && (src_h != net->resize_sh || src_w != net->resize_sw)) {
        CK(dnnb_sync_streams(net));          // ... or the old tables
        std::vector<int> tab(2 * (size_t)net->W + 2 * (size_t)net->H);
        resize_axis_table(src_w, net->W, true, tab.data(), tab.data() + net->W);
        resize_axis_table(src_h, net->H, false, tab.data() + 2 * net->W, tab.data() + 2 * net->W + net->H);
        if (!net->d_resize_tab) CK(cudaMalloc(&net->d_resize_tab, tab.size() * sizeof(int)));
        CK(cudaMemcpy(net->d_resize_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
        net->resize_sh = src_h;
        net->resize_sw = src_w;
    }
    return 0;
}
// uint8 frames of the network's own size go straight into the conv0 kernel when it is block 0
bool u8_fused(const dnnb_net *net, int src_h, int src_w) {
    const bool same = (src_h <= 0 || src_h == net->H) && (src_w <= 0 || src_w == net->W);
    return !net->knobs.no_u8_fusion && same && net->blocks.front().family == DNNB_KERNEL_UMMA_CONV0 && net->W % 16 == 0;
}

cudaError_t launch_preprocess(dnnb_net *net, const uint8_t *dev_u8, float *dev_f32, int swap_rb, cudaStream_t st,
                              int src_h = 0, int src_w = 0) {
    if (src_h > 0 && src_w > 0 && (src_h != net->H || src_w != net->W)) {
        simt::ResizeParams q{};
        q.src = dev_u8; q.dst = dev_f32;
        q.xofs = net->d_resize_tab; q.xw = q.xofs + net->W; q.yofs = q.xw + net->W; q.yw = q.yofs + net->H;
        q.lut = net->d_lut;
        q.B = net->B; q.SH = src_h; q.SW = src_w; q.DH = net->H; q.DW = net->W; q.swap_rb = swap_rb;
        simt::resize_bilinear_u8_kernel<<<ew_grid((size_t)net->B * net->H * net->W, net->num_sms), 256, 0, st>>>(q);
        return cudaGetLastError();
    }
    const size_t groups = (size_t)net->B * net->H * net->W / 4;
    simt::preprocess_u8_kernel<<<ew_grid(groups, net->num_sms), 256, 0, st>>>(
        reinterpret_cast<const uint32_t *>(dev_u8), reinterpret_cast<float4 *>(dev_f32), groups, net->d_lut, swap_rb);
    return cudaGetLastError();
}
}  // namespace

int dnnb_net_forward_u8_device(dnnb_net *net, const void *dev_u8, int bgr, void *dev_out, void *stream) {
    if (check_net(net, true)) return 1;
    if (!dev_u8) return fail("null device input");
    DeviceGuard g(net->device);
    if (ensure_u8(net)) return 1;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : net->stream;
    const bool fused = u8_fused(net, 0, 0);
    if (!fused) CK(launch_preprocess(net, static_cast<const uint8_t *>(dev_u8), net->d_in2[0], bgr ? 1 : 0, st));
    net->last_final = dev_out ? nullptr : static_cast<const float *>(net->blocks.back().d_out);
    net->last_in = fused ? nullptr : net->d_in2[0];
    return fused ? forward_impl(net, dev_u8, dev_out, st, true, bgr ? 1 : 0) : forward_impl(net, net->d_in2[0], dev_out, st);
}

namespace {
// One pipelined host-side forward: H2D on copy_in, kernels on net->stream, D2H on copy_out, with
// double-buffered device input / output so that the copies of forwards i-1 and i+1 overlap the
// kernels of forward i.  The result that travels back is either the head tensor (host_out) or,
// with dets/counts, only the label boxes the on-device post-processing leaves.
struct HostJob {
    const void *host_in = nullptr;
    bool u8 = false;
    int bgr = 0;
    int src_h = 0, src_w = 0;     // uint8 frames of another size: resized on the device (0 = network size)
    float *host_out = nullptr;
    dnnb_det *dets = nullptr;
    int *counts = nullptr;
};

int ensure_detect(dnnb_net *net) {
    const Block &l = net->blocks.back();
    if (l.oh != postproc::kGrid || l.ow != postproc::kGrid || l.oc != postproc::kBoxes * postproc::kFields)
        return fail("detection decoding needs a [B,13,13,125] network output, this one is [B,%d,%d,%d]", l.oh, l.ow, l.oc);
    if (!net->d_dets[0]) {
        for (int k = 0; k < 2; ++k) {
            CK(cudaMalloc(&net->d_dets[k], (size_t)net->B * net->max_det * sizeof(postproc::Det)));
            CK(cudaMalloc(&net->d_counts[k], (size_t)net->B * sizeof(int)));
        }
    }
    return 0;
}

cudaError_t launch_detect(dnnb_net *net, const float *pred, void *dets, void *counts, cudaStream_t st) {
    postproc::Params q{};
    q.pred = pred;
    q.dets = static_cast<postproc::Det *>(dets);
    q.counts = static_cast<int *>(counts);
    q.B = net->B;
    q.max_det = net->max_det;
    q.score_thr = net->score_thr;
    q.iou_thr = net->iou_thr;
    const int grid = net->B < net->num_sms * 4 ? net->B : net->num_sms * 4;
    postproc::detect_kernel<<<grid, postproc::kThreads, 0, st>>>(q);
    return cudaGetLastError();
}

int enqueue_host_forward(dnnb_net *net, const HostJob &job) {
    if (check_net(net, true)) return 1;
    const bool detect = job.dets != nullptr;
    if (!job.host_in || (!detect && !job.host_out) || (detect && !job.counts)) return fail("null host buffer");
    DeviceGuard g(net->device);
    if (job.u8 && (job.src_h < 0 || job.src_w < 0 || (job.src_h > 0) != (job.src_w > 0))) return fail("bad frame size");
    if (job.u8 && ensure_u8(net, job.src_h, job.src_w)) return 1;
    if (detect && ensure_detect(net)) return 1;
    const size_t n_in = job.u8 ? (size_t)net->B * (job.src_h ? job.src_h : net->H) * (job.src_w ? job.src_w : net->W) * 3
                               : (size_t)net->B * net->H * net->W * net->C;
    Block &l = net->blocks.back();
    if (!net->d_outf[0]) {
        net->d_outf[0] = static_cast<float *>(l.d_out);
        CK(cudaMalloc(&net->d_outf[1], l.out_elems * 4));
    }
    const int k = net->flip;
    net->flip ^= 1;
    // H2D on its own stream, as soon as the forward that last read this input buffer is done
    CK(cudaStreamWaitEvent(net->copy_in, net->in_free[k], 0));
    if (job.u8) CK(cudaMemcpyAsync(net->d_u8[k], job.host_in, n_in, cudaMemcpyHostToDevice, net->copy_in));
    else CK(cudaMemcpyAsync(net->d_in2[k], job.host_in, n_in * 4, cudaMemcpyHostToDevice, net->copy_in));
    CK(cudaEventRecord(net->in_ready[k], net->copy_in));
    // kernels: need the input, and the output buffer's previous D2H must have drained
    CK(cudaStreamWaitEvent(net->stream, net->in_ready[k], 0));
    CK(cudaStreamWaitEvent(net->stream, net->out_free[k], 0));
    const bool fused = job.u8 && u8_fused(net, job.src_h, job.src_w);
    if (job.u8 && !fused) CK(launch_preprocess(net, net->d_u8[k], net->d_in2[k], job.bgr ? 1 : 0, net->stream, job.src_h, job.src_w));
    if (fused ? forward_impl(net, net->d_u8[k], net->d_outf[k], net->stream, true, job.bgr ? 1 : 0)
              : forward_impl(net, net->d_in2[k], net->d_outf[k], net->stream))
        return 1;
    if (detect) CK(launch_detect(net, net->d_outf[k], net->d_dets[k], net->d_counts[k], net->stream));
    CK(cudaEventRecord(net->in_free[k], net->stream));
    CK(cudaEventRecord(net->out_ready[k], net->stream));
    // D2H on a third stream
    CK(cudaStreamWaitEvent(net->copy_out, net->out_ready[k], 0));
    if (detect) {
        CK(cudaMemcpyAsync(job.dets, net->d_dets[k], (size_t)net->B * net->max_det * sizeof(postproc::Det),
                           cudaMemcpyDeviceToHost, net->copy_out));
        CK(cudaMemcpyAsync(job.counts, net->d_counts[k], (size_t)net->B * sizeof(int), cudaMemcpyDeviceToHost, net->copy_out));
    } else {
        CK(cudaMemcpyAsync(job.host_out, net->d_outf[k], l.out_elems * 4, cudaMemcpyDeviceToHost, net->copy_out));
    }
    CK(cudaEventRecord(net->out_free[k], net->copy_out));
    net->last_final = net->d_outf[k];
    net->last_in = fused ? nullptr : net->d_in2[k];
    return 0;
}
}  // namespace

int dnnb_net_forward_async(dnnb_net *net, const float *host_in, float *host_out) {
    HostJob job;
    job.host_in = host_in;
    job.host_out = host_out;
    return enqueue_host_forward(net, job);
}

int dnnb_net_forward_u8_async(dnnb_net *net, const unsigned char *host_u8, int bgr, float *host_out) {
    HostJob job;
    job.host_in = host_u8;
    job.u8 = true;
    job.bgr = bgr;
    job.host_out = host_out;
    return enqueue_host_forward(net, job);
}

int dnnb_net_forward_frames_async(dnnb_net *net, const unsigned char *host_frames, int src_h, int src_w, int bgr,
                                  float *host_out) {
    HostJob job;
    job.host_in = host_frames;
    job.u8 = true;
    job.bgr = bgr;
    job.src_h = src_h;
    job.src_w = src_w;
    job.host_out = host_out;
    return enqueue_host_forward(net, job);
}

int dnnb_net_detect_frames_async(dnnb_net *net, const unsigned char *host_frames, int src_h, int src_w, int bgr,
                                 dnnb_det *dets, int *counts) {
    HostJob job;
    job.host_in = host_frames;
    job.u8 = true;
    job.bgr = bgr;
    job.src_h = src_h;
    job.src_w = src_w;
    job.dets = dets;
    job.counts = counts;
    return enqueue_host_forward(net, job);
}

int dnnb_net_set_detect(dnnb_net *net, float score_threshold, double iou_threshold, int max_det) {
    if (check_net(net, false)) return 1;
    if (max_det < 1 || max_det > postproc::kCand) return fail("max_det must be in [1, %d]", postproc::kCand);
    if (net->d_dets[0] && max_det != net->max_det) return fail("max_det cannot change after the first detect call");
    net->score_thr = score_threshold;
    net->iou_thr = iou_threshold;
    net->max_det = max_det;
    return 0;
}

int dnnb_net_detect_async(dnnb_net *net, const float *host_in, dnnb_det *dets, int *counts) {
    HostJob job;
    job.host_in = host_in;
    job.dets = dets;
    job.counts = counts;
    return enqueue_host_forward(net, job);
}

int dnnb_net_detect_u8_async(dnnb_net *net, const unsigned char *host_u8, int bgr, dnnb_det *dets, int *counts) {
    HostJob job;
    job.host_in = host_u8;
    job.u8 = true;
    job.bgr = bgr;
    job.dets = dets;
    job.counts = counts;
    return enqueue_host_forward(net, job);
}

int dnnb_net_detect_device(dnnb_net *net, const void *dev_in, int is_u8, int bgr, void *dev_dets, void *dev_counts,
                           void *stream) {
    if (check_net(net, true)) return 1;
    if (!dev_in || !dev_dets || !dev_counts) return fail("null device buffer");
    DeviceGuard g(net->device);
    if (ensure_detect(net)) return 1;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : net->stream;
    const void *in = dev_in;
    bool fused = false;
    if (is_u8) {
        if (ensure_u8(net)) return 1;
        fused = u8_fused(net, 0, 0);
        if (!fused) {
            CK(launch_preprocess(net, static_cast<const uint8_t *>(dev_in), net->d_in2[0], bgr ? 1 : 0, st));
            in = net->d_in2[0];
        }
    }
    net->last_final = static_cast<const float *>(net->blocks.back().d_out);
    net->last_in = fused ? nullptr : in;
    if (forward_impl(net, in, nullptr, st, fused, bgr ? 1 : 0)) return 1;
    CK(launch_detect(net, net->last_final, dev_dets, dev_counts, st));
    return 0;
}

int dnnb_net_sync(dnnb_net *net) {
    if (check_net(net, false)) return 1;
    DeviceGuard g(net->device);
    CK(cudaStreamSynchronize(net->stream));
    if (net->copy_in) CK(cudaStreamSynchronize(net->copy_in));
    if (net->copy_out) CK(cudaStreamSynchronize(net->copy_out));
    return 0;
}

int dnnb_net_forward(dnnb_net *net, const float *host_in, float *host_out) {
    if (dnnb_net_forward_async(net, host_in, host_out)) return 1;
    return dnnb_net_sync(net);
}

int dnnb_net_block_output(dnnb_net *net, int block, float *host_out) {
    if (check_net(net, true)) return 1;
    if (block < 0 || block >= (int)net->blocks.size() || !host_out) return fail("bad block index or buffer");
    DeviceGuard g(net->device);
    Block &b = net->blocks[block];
    CK(cudaStreamSynchronize(net->stream));
    if (!b.out_bf16) {
        const void *src = b.d_out;
        if (block + 1 == (int)net->blocks.size() && net->last_final) src = net->last_final;
        CK(cudaMemcpy(host_out, src, b.out_elems * 4, cudaMemcpyDeviceToHost));
    } else {
        DevBuf tmp;
        if (tmp.alloc(b.out_elems * 4)) return 1;
        simt::cast_output_kernel<__nv_bfloat16><<<ew_grid(b.out_elems, net->num_sms), 256>>>(
            (const __nv_bfloat16 *)b.d_out, tmp.f(), b.out_elems);
        CK(cudaGetLastError());
        CK(cudaMemcpy(host_out, tmp.p, b.out_elems * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int dnnb_net_debug_prepool(dnnb_net *net, int block, float *host_out) {
    if (check_net(net, true)) return 1;
    if (block < 0 || block >= (int)net->blocks.size() || !host_out) return fail("bad block index or buffer");
    Block &b = net->blocks[block];
    if (!b.pool_fused || b.family == DNNB_KERNEL_SIMT) return fail("block %d has no tcgen05-fused pool", block);
    DeviceGuard g(net->device);
    const size_t n = (size_t)net->B * b.ch * b.cw * b.oc;
    DevBuf tmp;
    if (tmp.alloc(n * 4)) return 1;
    const void *in = block == 0 ? (net->d_in_cast ? net->d_in_cast : (net->last_in ? net->last_in : (const void *)net->d_in))
                                : net->blocks[block - 1].d_out;
    CK(cudaStreamSynchronize(net->stream));
    if (run_block(net, b, in, nullptr, net->stream, tmp.f())) return 1;
    CK(cudaStreamSynchronize(net->stream));
    CK(cudaMemcpy(host_out, tmp.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int dnnb_net_block_info(dnnb_net *net, int block, int *info) {
    if (check_net(net, true)) return 1;
    if (block < 0 || block >= (int)net->blocks.size() || !info) return fail("bad block index or buffer");
    const Block &b = net->blocks[block];
    info[0] = b.pair ? DNNB_KERNEL_UMMA_PAIR : b.family; info[1] = b.oh; info[2] = b.ow; info[3] = b.oc;
    info[4] = b.pool_fused ? 1 : 0;
    info[5] = b.bn; info[6] = b.swz;
    info[7] = b.pair ? b.pair->stages : b.umma ? b.umma->stages : b.rowpair ? b.rowpair->stages : b.family == DNNB_KERNEL_UMMA_CONV0 ? conv0::kStages : 0;
    info[8] = b.grid;
    info[9] = b.pool_separate ? 1 : 0;
    info[10] = b.out_bf16 ? 2 : 4;
    info[11] = 1 + (b.pool_separate ? 1 : 0) + (b.umma_tail ? 1 : 0) + (b.pair_tail ? 1 : 0);
    return 0;
}

int dnnb_net_time_blocks(dnnb_net *net, int warmup, int iters, float *ms) {
    if (check_net(net, true)) return 1;
    if (!ms || iters <= 0 || warmup < 0) return fail("bad argument");
    DeviceGuard g(net->device);
    const size_t flush_bytes = 256u << 20;   // > 126 MB L2
    if (!net->d_flush) CK(cudaMalloc(&net->d_flush, flush_bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    cudaStream_t st = net->stream;
    // one full forward first so every block has a real input
    if (forward_impl(net, net->d_in, nullptr, st)) return 1;
    const void *cur = net->d_in_cast ? net->d_in_cast : (const void *)net->d_in;
    for (size_t i = 0; i < net->blocks.size(); ++i) {
        Block &b = net->blocks[i];
        double total = 0.0;
        for (int it = 0; it < warmup + iters; ++it) {
            CK(cudaMemsetAsync(net->d_flush, it & 0xff, flush_bytes, st));
            CK(cudaEventRecord(e0, st));
            if (run_block(net, b, cur, nullptr, st)) return 1;
            CK(cudaEventRecord(e1, st));
            CK(cudaEventSynchronize(e1));
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (it >= warmup) total += t;
        }
        ms[i] = (float)(total / iters);
        cur = b.d_out;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

int dnnb_net_profile_block(dnnb_net *net, int block) {
    if (check_net(net, true)) return 1;
    if (block >= (int)net->blocks.size()) return fail("bad block index");
    net->prof_block = block;   // < 0 switches profiling off
    net->prof_used = 0;
    return 0;
}

int dnnb_net_profile_read(dnnb_net *net, float *total_ms, int *count) {
    if (check_net(net, true) || !total_ms || !count) return fail("bad argument");
    DeviceGuard g(net->device);
    double total = 0.0;
    for (size_t i = 0; i < net->prof_used; ++i) {
        CK(cudaEventSynchronize(net->prof_events[i].second));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, net->prof_events[i].first, net->prof_events[i].second));
        total += t;
    }
    *total_ms = (float)total;
    *count = (int)net->prof_used;
    net->prof_used = 0;
    return 0;
}

int dnnb_net_launches_per_forward(dnnb_net *net, int *n) {
    if (check_net(net, true) || !n) return fail("bad argument");
    *n = net->launches;
    return 0;
}

int dnnb_net_debug_block_limit(dnnb_net *net, int blocks) {
    if (check_net(net, true)) return 1;
    DeviceGuard g(net->device);
    CK(dnnb_sync_streams(net));
    net->block_limit = blocks;
    for (dnnb_net::GraphSlot &gs : net->graphs) cudaGraphExecDestroy(gs.exec);   // recorded for another limit
    net->graphs.clear();
    net->warm_in = nullptr;
    net->warm_out = nullptr;
    return 0;
}

int dnnb_net_destroy(dnnb_net *net) {
    if (!net) return 0;
    DeviceGuard g(net->device);
    if (net->stream) cudaStreamSynchronize(net->stream);
    for (Block &b : net->blocks) {
        if (b.pool_separate && b.d_conv_out) cudaFree(b.d_conv_out);
        if (b.d_out) cudaFree(b.d_out);
        if (b.d_kernel) cudaFree(b.d_kernel);
        if (b.d_wpack) cudaFree(b.d_wpack);
        if (b.d_workspace) cudaFree(b.d_workspace);
        if (b.d_counters) cudaFree(b.d_counters);
        if (b.d_scale) cudaFree(b.d_scale);
        if (b.d_shift) cudaFree(b.d_shift);
    }
    for (auto &pr : net->prof_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (dnnb_net::GraphSlot &g : net->graphs) cudaGraphExecDestroy(g.exec);
    if (net->d_in2[0]) cudaFree(net->d_in2[0]);
    if (net->d_in2[1]) cudaFree(net->d_in2[1]);
    if (net->d_outf[1]) cudaFree(net->d_outf[1]);
    if (net->d_u8[0]) cudaFree(net->d_u8[0]);
    if (net->d_u8[1]) cudaFree(net->d_u8[1]);
    if (net->d_lut) cudaFree(net->d_lut);
    if (net->d_resize_tab) cudaFree(net->d_resize_tab);
    if (net->d_lut_conv0) cudaFree(net->d_lut_conv0);
    for (int k = 0; k < 2; ++k) {
        if (net->d_dets[k]) cudaFree(net->d_dets[k]);
        if (net->d_counts[k]) cudaFree(net->d_counts[k]);
        if (net->in_ready[k]) cudaEventDestroy(net->in_ready[k]);
        if (net->in_free[k]) cudaEventDestroy(net->in_free[k]);
        if (net->out_ready[k]) cudaEventDestroy(net->out_ready[k]);
        if (net->out_free[k]) cudaEventDestroy(net->out_free[k]);
    }
    if (net->copy_in) cudaStreamDestroy(net->copy_in);
    if (net->copy_out) cudaStreamDestroy(net->copy_out);
    if (net->d_in_cast) cudaFree(net->d_in_cast);
    if (net->d_flush) cudaFree(net->d_flush);
    if (net->stream) cudaStreamDestroy(net->stream);
    delete net;
    return 0;
}

// ---------------------------------------------------------------------------- input preprocessing
int dnnb_resize_input(int device, const unsigned char *frames_host, int batch, int src_h, int src_w, int dst_h,
                      int dst_w, int bgr, float *out_host) {
    if (!frames_host || !out_host) return fail("dnnb_resize_input: null pointer");
    if (batch < 1 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1) return fail("dnnb_resize_input: bad shape");
    DeviceGuard g(device);
    const size_t n_src = (size_t)batch * src_h * src_w * 3, n_dst = (size_t)batch * dst_h * dst_w * 3;
    std::vector<int> tab(2 * (size_t)dst_w + 2 * (size_t)dst_h);
    resize_axis_table(src_w, dst_w, true, tab.data(), tab.data() + dst_w);
    resize_axis_table(src_h, dst_h, false, tab.data() + 2 * dst_w, tab.data() + 2 * dst_w + dst_h);
    float lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = (float)((double)i / 255.0);
    DevBuf dsrc, ddst, dtab, dlut;
    if (dsrc.alloc(n_src) || ddst.alloc(n_dst * 4) || dtab.alloc(tab.size() * 4) || dlut.alloc(sizeof(lut))) return 1;
    CK(cudaMemcpy(dsrc.p, frames_host, n_src, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dtab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dlut.p, lut, sizeof(lut), cudaMemcpyHostToDevice));
    simt::ResizeParams q{};
    q.src = static_cast<const uint8_t *>(dsrc.p); q.dst = ddst.f();
    q.xofs = static_cast<const int *>(dtab.p); q.xw = q.xofs + dst_w; q.yofs = q.xw + dst_w; q.yw = q.yofs + dst_h;
    q.lut = dlut.f();
    q.B = batch; q.SH = src_h; q.SW = src_w; q.DH = dst_h; q.DW = dst_w; q.swap_rb = bgr ? 1 : 0;
    simt::resize_bilinear_u8_kernel<<<ew_grid((size_t)batch * dst_h * dst_w, sm_count()), 256>>>(q);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out_host, ddst.p, n_dst * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// ---------------------------------------------------------------------------- detection decoding
// Stand-alone form of the on-device post-processing: host head tensors in, label boxes out.
namespace {
int postprocess_host(int device, const float *pred_host, const dnnb_det *cand_host, int batch, float score_thr,
                     double iou_thr, int max_det, dnnb_det *cand_dump_host, dnnb_det *dets_host, int *counts_host) {
    if (batch < 1) return fail("batch must be positive");
    if (max_det < 1 || max_det > postproc::kCand) return fail("max_det must be in [1, %d]", postproc::kCand);
    DeviceGuard g(device);
    static_assert(sizeof(dnnb_det) == sizeof(postproc::Det), "dnnb_det layout");
    const size_t n_pred = (size_t)batch * postproc::kCand * postproc::kFields;
    const size_t n_cand = (size_t)batch * postproc::kCand;
    DevBuf dpred, dcand, ddump, ddets, dcounts;
    postproc::Params q{};
    if (pred_host) {
        if (dpred.alloc(n_pred * 4)) return 1;
        CK(cudaMemcpy(dpred.p, pred_host, n_pred * 4, cudaMemcpyHostToDevice));
        q.pred = dpred.f();
    }
    if (cand_host) {
        if (dcand.alloc(n_cand * sizeof(dnnb_det))) return 1;
        CK(cudaMemcpy(dcand.p, cand_host, n_cand * sizeof(dnnb_det), cudaMemcpyHostToDevice));
        q.cand_in = static_cast<const postproc::Det *>(dcand.p);
    }
    if (cand_dump_host) {
        if (ddump.alloc(n_cand * sizeof(dnnb_det))) return 1;
        q.cand_out = static_cast<postproc::Det *>(ddump.p);
    }
    if (ddets.alloc((size_t)batch * max_det * sizeof(dnnb_det)) || dcounts.alloc((size_t)batch * sizeof(int))) return 1;
    CK(cudaMemset(ddets.p, 0, (size_t)batch * max_det * sizeof(dnnb_det)));
    q.dets = static_cast<postproc::Det *>(ddets.p);
    q.counts = static_cast<int *>(dcounts.p);
    q.B = batch;
    q.max_det = max_det;
    q.score_thr = score_thr;
    q.iou_thr = iou_thr;
    const int sms = sm_count();
    postproc::detect_kernel<<<batch < sms * 4 ? batch : sms * 4, postproc::kThreads>>>(q);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (cand_dump_host) CK(cudaMemcpy(cand_dump_host, ddump.p, n_cand * sizeof(dnnb_det), cudaMemcpyDeviceToHost));
    if (dets_host) CK(cudaMemcpy(dets_host, ddets.p, (size_t)batch * max_det * sizeof(dnnb_det), cudaMemcpyDeviceToHost));
    if (counts_host) CK(cudaMemcpy(counts_host, dcounts.p, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}
}  // namespace

int dnnb_postprocess(int device, const float *pred_host, int batch, float score_threshold, double iou_threshold,
                     int max_det, dnnb_det *dets_host, int *counts_host) {
    if (!pred_host || !dets_host || !counts_host) return fail("dnnb_postprocess: null pointer");
    return postprocess_host(device, pred_host, nullptr, batch, score_threshold, iou_threshold, max_det, nullptr,
                            dets_host, counts_host);
}

int dnnb_postprocess_candidates(int device, const float *pred_host, int batch, dnnb_det *cand_host) {
    if (!pred_host || !cand_host) return fail("dnnb_postprocess_candidates: null pointer");
    return postprocess_host(device, pred_host, nullptr, batch, 0.3f, 0.3, 1, cand_host, nullptr, nullptr);
}

int dnnb_nms(int device, const dnnb_det *cand_host, int batch, float score_threshold, double iou_threshold, int max_det,
             dnnb_det *dets_host, int *counts_host) {
    if (!cand_host || !dets_host || !counts_host) return fail("dnnb_nms: null pointer");
    return postprocess_host(device, nullptr, cand_host, batch, score_threshold, iou_threshold, max_det, nullptr,
                            dets_host, counts_host);
}

// pinned host buffers for the end-to-end path (async H2D / D2H need page-locked memory)
int dnnb_host_alloc(size_t bytes, void **out) {
    if (!out) return fail("null out pointer");
    CK(cudaHostAlloc(out, bytes ? bytes : 4, cudaHostAllocPortable));
    return 0;
}
int dnnb_host_free(void *p) {
    if (p) CK(cudaFreeHost(p));
    return 0;
}

// seconds <= 0: burst (best of 5 short launches); otherwise back-to-back launches for `seconds`
// and the mean rate of the second half (the clock the part settles at under its power cap)
static int measure_mma_peak_impl(int device, int precision, float seconds, float *tflops) {
    if (!tflops) return fail("null argument");
    if (precision != DNNB_TF32 && precision != DNNB_BF16) return fail("precision must be TF32 or BF16");
    DeviceGuard g(device);
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int smem = 128 * 128 + 256 * 128 + 1024;
    const int iters = 4096;
    auto run = [&](int n) -> cudaError_t {
        if (precision == DNNB_BF16) {
            cudaFuncSetAttribute(umma_conv::mma_peak_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            umma_conv::mma_peak_kernel<true><<<sms, 64, smem>>>(n, seconds > 0.f ? 1 : 0);
        } else {
            cudaFuncSetAttribute(umma_conv::mma_peak_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            umma_conv::mma_peak_kernel<false><<<sms, 64, smem>>>(n, seconds > 0.f ? 1 : 0);
        }
        return cudaGetLastError();
    };
    const double k_per_mma = precision == DNNB_BF16 ? 16.0 : 8.0;
    const double flops = (double)sms * iters * 4.0 * 2.0 * 128.0 * 256.0 * k_per_mma;
    CK(run(256));
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    if (seconds <= 0.f) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            CK(run(iters));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (t < best) best = t;
        }
        *tflops = (float)(flops / (best * 1e-3) / 1e12);
    } else {
        std::vector<float> times;
        double elapsed = 0.0;
        while (elapsed < seconds * 1e3 && times.size() < 100000) {
            CK(cudaEventRecord(e0));
            for (int k = 0; k < 8; ++k) CK(run(iters));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, e0, e1));
            times.push_back(t / 8.f);
            elapsed += t;
        }
        double sum = 0.0;
        size_t n = 0;
        for (size_t i = times.size() / 2; i < times.size(); ++i) { sum += times[i]; ++n; }
        *tflops = (float)(flops / (sum / n * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

int dnnb_measure_mma_peak(int device, int precision, float *tflops) {
    return measure_mma_peak_impl(device, precision, 0.f, tflops);
}

int dnnb_measure_mma_peak_sustained(int device, int precision, float seconds, float *tflops) {
    if (!(seconds > 0.f) || seconds > 30.f) return fail("seconds must be in (0, 30]");
    return measure_mma_peak_impl(device, precision, seconds, tflops);
}

}  // extern "C"
