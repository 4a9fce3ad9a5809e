// planner.cuh -- the resident network's state and its planner: the environment switches (Knobs), the
// per-block description and plan (Block), the network handle (dnnb_net), the tensor-map encoders for
// activations and weights, and plan_block(), which decides for every conv block of
// proj3/yolov2tiny.py:30-77 which kernel runs it, with which tiles, and how its tail rows are cut.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dnn_b200.h"
#include "conv0_umma.cuh"
#include "launchers.cuh"

extern "C" char **environ;

namespace {

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
    bool pair_no_tail = false;       // DNNB_PAIR_NO_TAIL     pairs also run the partial last wave (no tail launches at all)
    bool pair_force_tail = false;    // DNNB_PAIR_FORCE_TAIL  tail launches also behind short K loops (tests)
    bool no_pair_tail = false;       // DNNB_NO_PAIR_TAIL     rows past the full waves go to the single-CTA kernel only
    bool no_tail = false;            // DNNB_NO_TAIL          no main/tail cut of wave-quantised 256 x 256 blocks
    int lat_fuse_nb = 0;             // DNNB_LAT_FUSE_NB      latency plan: keep fused spatial tiles up to this many images
    int split_min_stages = 0;        // DNNB_SPLIT_MIN_STAGES latency plan: K blocks per split (0: default 4, short loops whole)
    bool no_u8_fusion = false;       // DNNB_NO_U8_FUSION     uint8 frames through a separate preprocessing kernel
    bool no_pool1_fusion = false;    // DNNB_NO_POOL1_FUSION  block 5: the stride-1 SAME pool as a separate kernel
    bool no_cluster_splitk = false;  // DNNB_NO_CLUSTER_SPLITK latency plan: split-K through the global workspace, not cluster DSMEM
    bool no_tma_store = false;       // DNNB_NO_TMA_STORE     epilogues store with per-thread st.global instead of bulk copies
    bool pdl_all = false;            // DNNB_PDL_ALL          programmatic dependent launch in the default (throughput) plan too
    int lat_bn = 0;                  // DNNB_LAT_BN=n         latency plan: cap on the N tile for every block (default: 128 where it keeps the pools fused)
    bool verbose_plan = false;       // DNNB_VERBOSE_PLAN     print planning decisions to stderr
};

// Every DNNB_* variable the library understands.  Anything else that starts with DNNB_ is a typo or a stale
// switch from another build and makes dnnb_net_finalize fail (a stray variable must not silently change the
// plan a benchmark measures).  Flags: present (empty or "1") = on, "0" = off.  Numbers: decimal, within [lo, hi].
struct KnobDef {
    const char *name;
    bool is_flag;
    int lo, hi;
};
const KnobDef g_knob_defs[] = {
    {"DNNB_LATENCY_MODE", true, 0, 1},     {"DNNB_NO_PDL", true, 0, 1},          {"DNNB_NO_GRAPH", true, 0, 1},
    {"DNNB_FORCE_SIMT", true, 0, 1},       {"DNNB_NO_CONV0", true, 0, 1},        
    {"DNNB_NO_ROWPAIR", true, 0, 1},       {"DNNB_ROWPAIR_LINEAR", true, 0, 1},  {"DNNB_NO_ROWPAIR_LINEAR", true, 0, 1},
    {"DNNB_ROWPAIR_MAX_TR", false, 2, 8},  {"DNNB_NO_POOL_FIRST", true, 0, 1},   {"DNNB_NO_STAGE_OUT", true, 0, 1},
    {"DNNB_NO_MT2", true, 0, 1},           {"DNNB_MT2_MASK", false, 0, 1 << 20}, {"DNNB_NO_PAIR", true, 0, 1},
    {"DNNB_PAIR_MT", false, 1, 2},         {"DNNB_PAIR_MIN_TILES", false, 0, 1 << 30}, {"DNNB_PAIR_NO_TAIL", true, 0, 1}, {"DNNB_PAIR_FORCE_TAIL", true, 0, 1},
    {"DNNB_NO_PAIR_TAIL", true, 0, 1},     {"DNNB_NO_TAIL", true, 0, 1},         {"DNNB_LAT_FUSE_NB", false, 0, 128},
    {"DNNB_SPLIT_MIN_STAGES", false, 0, 1024}, {"DNNB_NO_U8_FUSION", true, 0, 1}, {"DNNB_NO_POOL1_FUSION", true, 0, 1},
    {"DNNB_NO_CLUSTER_SPLITK", true, 0, 1}, {"DNNB_NO_TMA_STORE", true, 0, 1},   {"DNNB_VERBOSE_PLAN", true, 0, 1}, {"DNNB_LAT_BN", false, 0, 256}, {"DNNB_PDL_ALL", true, 0, 1},
};

// Reads the switches.  `active` receives "NAME=value" for every switch that is set to something other than
// its default, comma separated (dnnb_net_knobs reports it; bench.py records it and refuses to publish a
// headline number with any of them set).  Returns non-zero with `err` filled for an unknown DNNB_* name or
// a value that does not parse.
int read_knobs(Knobs &k, std::string &active, std::string &err) {
    active.clear();
    for (char **e = environ; e && *e; ++e) {
        if (strncmp(*e, "DNNB_", 5) != 0) continue;
        const char *eq = strchr(*e, '=');
        const std::string name(*e, eq ? (size_t)(eq - *e) : strlen(*e));
        bool known = false;
        for (const KnobDef &d : g_knob_defs) known = known || name == d.name;
        if (!known) {
            err = "unknown environment switch " + name + " (include/dnn_b200.h lists the DNNB_* variables; unset it)";
            return 1;
        }
    }
    bool bad = false;
    auto value = [&](const char *name, int dflt) -> int {
        const KnobDef *def = nullptr;
        for (const KnobDef &d : g_knob_defs)
            if (strcmp(d.name, name) == 0) def = &d;
        const char *v = getenv(name);
        if (!v || !def) return dflt;
        long x;
        if (def->is_flag && v[0] == 0) x = 1;
        else {
            char *end = nullptr;
            x = strtol(v, &end, 10);
            if (end == v || *end != 0 || x < def->lo || x > def->hi) {
                err = std::string("bad value for ") + name + ": '" + v + "'";
                bad = true;
                return dflt;
            }
        }
        if ((int)x != dflt) {
            if (!active.empty()) active += ",";
            active += std::string(name) + "=" + std::to_string(x);
        }
        return (int)x;
    };
    k = Knobs();
    k.latency = value("DNNB_LATENCY_MODE", 0) != 0;
    k.no_pdl = value("DNNB_NO_PDL", 0) != 0;
    k.no_graph = value("DNNB_NO_GRAPH", 0) != 0;
    k.force_simt = value("DNNB_FORCE_SIMT", 0) != 0;
    k.no_conv0 = value("DNNB_NO_CONV0", 0) != 0;
    k.no_rowpair = value("DNNB_NO_ROWPAIR", 0) != 0;
    const int lin_on = getenv("DNNB_ROWPAIR_LINEAR") ? value("DNNB_ROWPAIR_LINEAR", -1) : -1;
    k.rowpair_linear = lin_on >= 0 ? lin_on : (value("DNNB_NO_ROWPAIR_LINEAR", 0) != 0 ? 0 : -1);
    k.rowpair_max_tr = value("DNNB_ROWPAIR_MAX_TR", 8);
    k.no_pool_first = value("DNNB_NO_POOL_FIRST", 0) != 0;
    k.no_stage_out = value("DNNB_NO_STAGE_OUT", 0) != 0;
    k.no_mt2 = value("DNNB_NO_MT2", 0) != 0;
    k.mt2_mask = getenv("DNNB_MT2_MASK") ? value("DNNB_MT2_MASK", -1) : -1;
    k.no_pair = value("DNNB_NO_PAIR", 0) != 0;
    k.pair_mt = value("DNNB_PAIR_MT", 1);
    k.pair_min_tiles = getenv("DNNB_PAIR_MIN_TILES") ? value("DNNB_PAIR_MIN_TILES", -1) : -1;
    k.pair_no_tail = value("DNNB_PAIR_NO_TAIL", 0) != 0;
    k.pair_force_tail = value("DNNB_PAIR_FORCE_TAIL", 0) != 0;
    k.no_pair_tail = value("DNNB_NO_PAIR_TAIL", 0) != 0;
    k.no_tail = value("DNNB_NO_TAIL", 0) != 0;
    k.lat_fuse_nb = value("DNNB_LAT_FUSE_NB", 0);
    k.split_min_stages = value("DNNB_SPLIT_MIN_STAGES", 0);
    k.no_u8_fusion = value("DNNB_NO_U8_FUSION", 0) != 0;
    k.no_pool1_fusion = value("DNNB_NO_POOL1_FUSION", 0) != 0;
    k.no_cluster_splitk = value("DNNB_NO_CLUSTER_SPLITK", 0) != 0;
    k.no_tma_store = value("DNNB_NO_TMA_STORE", 0) != 0;
    k.verbose_plan = value("DNNB_VERBOSE_PLAN", 0) != 0;
    k.lat_bn = value("DNNB_LAT_BN", 0);
    k.pdl_all = value("DNNB_PDL_ALL", 0) != 0;
    return bad ? 1 : 0;
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
    int lat_bn = 0;                             // latency plan: cap on the N tile (0 = none)
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
    unsigned long long *d_timeline = nullptr;   // debug: per-CTA phase stamps of the generic kernel (dnnb_net_debug_timeline)
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
    std::string knobs_active;            // "NAME=value,..." of the DNNB_* switches that are not at their default
    int latency_override = -1;           // dnnb_net_set_latency: -1 = DNNB_LATENCY_MODE decides
    std::vector<int> lat_widths;         // dnnb_net_set_latency_widths: per-block N-tile width under the latency plan (0 = modelled choice)
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

// Modelled time (us) of a planned generic-kernel block at small batch, for comparing N-tile widths under the latency
// plan (dnnb_net_finalize).  < 0: the block does not run the generic kernel (nothing to choose).  Fitted to the phase
// stamps of tools/batch1_timeline.py and the per-block times of tools/batch1_blocks.py at batch 1 (profiles/
// r02_SUMMARY.md); it only has to order the candidates of one block:
//  * K loop: a CTA takes 0.10 us + 1 us per 262 KB for every K block it streams (block 3: 18 blocks of 20 KB in 2.9 us;
//    block 7: 32 blocks of 32 KB in 7.1 us), and all CTAs together get 21 TB/s out of L2;
//  * partial sums through the global workspace: 5.7 us + 18 ns per column of the tile (store, flag, reload, sum);
//    through distributed shared memory: 2 us + 20 ns per column + a term that grows with the cluster (1.5 us at 8 CTAs:
//    such a kernel's CTAs start only once whole groups of SMs are free, its prologue no longer hides behind the
//    preceding kernel; 4 us at 16);
//  * unsplit: 11 ns per column of epilogue; a pool that is not fused: 6 us for its own kernel.
double latency_cost(const Block &b, int num_sms) {
    if ((b.family != DNNB_KERNEL_UMMA_TILED && b.family != DNNB_KERNEL_UMMA_IM2COL) || !b.umma || b.pair) return -1.0;
    const umma_conv::Params &p = b.up;
    const int es = b.in_bf16 ? 2 : 4;
    const int row_bytes = b.ic * es;
    const int swz = row_bytes >= 128 ? 128 : row_bytes;
    const int kblocks = b.kh * b.kw * (row_bytes / swz);                 // K blocks of one swizzle span
    const int splits = p.splits > 1 ? p.splits : 1;
    const int kb_cta = (kblocks + splits - 1) / splits;
    const double kb_bytes = (double)(b.umma->key.mt * umma_conv::kBM + b.bn) * swz / 1024.0;     // KB per K block
    const int ctas = p.num_m_tiles * p.num_n_tiles * splits;
    const int waves = (ctas + num_sms - 1) / num_sms;
    double k_time = kb_cta * (0.10 + kb_bytes / 262.0);
    const double shared = (double)(ctas < num_sms ? ctas : num_sms) * kb_cta * kb_bytes / 21000.0;
    if (shared > k_time) k_time = shared;
    k_time *= waves;
    double exchange;
    if (splits == 1) exchange = 0.011 * b.bn;
    else if (p.cluster_splitk) exchange = 2.0 + 0.02 * b.bn + (splits >= 16 ? 4.0 : splits >= 8 ? 1.5 : splits >= 4 ? 0.5 : 0.3);
    else exchange = 5.7 + 0.018 * b.bn;
    return k_time + exchange + (b.pool_separate ? 6.0 : 0.0);
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
    b.pool_separate = false;
    // a block may be planned more than once (BF16 block 0): start from a clean plan every time
    b.umma = nullptr; b.umma_tail = nullptr; b.pair = nullptr; b.pair_tail = nullptr; b.rowpair = nullptr;
    b.rp_linear = false; b.op_bf16 = false;
    b.grid = b.grid_tail = b.grid_pair_tail = 0;
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
        int bn = b.oc <= 32 ? 32 : b.oc <= 64 ? 64 : b.oc <= 128 ? 128 : 256;
        // latency plan: a narrower N tile means more tiles and fewer K slices per tile -- half the partial-sum bytes per CTA
        // in both reduction phases (dnnb_net_finalize tries 128 first and falls back when that would un-fuse a pool)
        if (kn.latency && b.lat_bn > 0 && bn > b.lat_bn) bn = b.lat_bn;
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
            // A pool fused into the IM2COL epilogue (epilogue_pool1.cuh), for blocks whose tiles are runs of consecutive
            // pixels: (1) the 2x2 stride-1 SAME pool of block 5 -- tiles advance by 128 - (W + 1) output pixels, the last
            // W + 1 rows of each tile are halo; (2) a 2x2/2 pool when the spatial tiling was given up (small batches under
            // the latency plan) -- tiles are whole pairs of image rows, 2kW <= 128 pixels.  Split-K is compatible only in
            // its cluster form (all rows of a column strip end up in one CTA), so the split is decided first.
            const int total_kstages = total_kb / kbs;
            auto decide_split = [&](const UmmaEntry *ent, int tiles_mn, int mt_, int &splits, int &cluster) {
                splits = 1;
                cluster = 0;
                // Splitting costs a few us (partials exchanged, sibling wait, reduction) and a CTA streams one ring of smem
                // stages per ~2 us round trip, so a K loop of up to three ring passes stays whole (measured, batch 1:
                // block 3, 18 K blocks on 6 stages, 21.5 us split four ways, 14.8 us whole).
                const bool long_k = total_kstages > 3 * ent->stages || kn.split_min_stages > 0;
                if (fuse || !(tiles_mn * 2 <= net->num_sms && total_kstages >= 8 && kn.latency)) return;
                int sp = net->num_sms / tiles_mn;
                const int min_stages = kn.split_min_stages > 0 ? kn.split_min_stages : 4;
                if (sp > total_kstages / min_stages) sp = total_kstages / min_stages;
                if (sp < 2) return;
                // Cluster split-K: the K slices of a tile are the CTAs of one thread-block cluster (2, 4, 8 or 16) and the
                // partial sums are reduced through distributed shared memory instead of a workspace in global memory
                // guarded by counters (conv_umma.cuh).  Needs every tile's cluster resident at once, which caps the slices
                // (8 tiles: clusters of 8 at most on this part), so the two forms are compared with a small cost model from
                // the measured phase times (tools/batch1_timeline.py, batch 1): a K block streams in ~1.7 ns per column of
                // the N tile; the cluster reduction takes ~1.5 us + the tile's bytes over distributed shared memory at
                // ~40 GB/s per SM, the workspace form ~7.5 (BN 128) / 9 us (BN 256) -- plus a pool kernel where one follows.
                // The cluster form also pays for K loops too short for the workspace form when the tiles alone would leave
                // most SMs idle (BF16: block 8, 2 tiles, 14.4 -> 8.1 us; but block 3, 26 tiles of 9 K blocks, 10.4 -> 14.9).
                const double t_kb = 0.0017 * bn;
                const double cost_ws = long_k ? (double)total_kstages / sp * t_kb + (bn > 128 ? 9.0 : 7.5) + (b.pool_k > 0 ? 5.0 : 0.0) : 1e9;
                if (!kn.no_cluster_splitk && mt_ == 1 && !ent->key.stage_out && ent->max_clusters && (long_k || tiles_mn * 8 <= net->num_sms)) {
                    int spc = 16;
                    while (spc > sp) spc /= 2;
                    if (kn.verbose_plan)
                        for (int c = 16; c >= 2; --c)
                            fprintf(stderr, "block %zu: tiles_mn %d, clusters of %d resident at once: %d\n", i, tiles_mn, c, ent->max_clusters(c));
                    while (spc >= 2 && ent->max_clusters(spc) < tiles_mn) spc /= 2;
                    const double cost_cl = spc >= 2 ? (double)total_kstages / spc * t_kb + 1.5 + 0.0128 * bn : 1e9;
                    if (kn.verbose_plan)
                        fprintf(stderr, "block %zu: BN %d, %d K stages: workspace x%d ~%.1f us, cluster x%d ~%.1f us\n", i, bn, total_kstages,
                                sp, cost_ws, spc, cost_cl);
                    if (spc >= 2 && cost_cl + 1.0 <= cost_ws) { splits = spc; cluster = 1; return; }      // a tie goes to the form with more CTAs
                }
                if (long_k) splits = sp;
            };
            int ipool = 0;           // 0 none, 1 = 2x2/1 SAME, 2 = 2x2/2
            int stride = umma_conv::kBM;
            if (!fuse && b.pool_k == 2 && swz == 128 && kbs == 1 && b.oc % 4 == 0 && !kn.no_pool1_fusion) {
                if (b.pool_s == 1 && b.pool_same && b.oh == b.ch && b.ow == b.cw && b.ppt == 0 && b.ppl == 0 && b.iw == pool1::kW) {
                    ipool = 1;
                    stride = umma_conv::kBM - (b.iw + 1);
                } else if (b.pool_s == 2 && b.ih % 2 == 0 && b.iw % 2 == 0 && 2 * b.iw <= umma_conv::kBM) {
                    ipool = 2;
                    stride = umma_conv::kBM / (2 * b.iw) * (2 * b.iw);
                }
            }
            UmmaKey key{in_bf16, b.out_bf16, bn, swz, kbs, !fuse, fuse, 1, false, ipool != 0};
            const UmmaEntry *e = find_umma(key);
            if (ipool) {
                int sp = 1, cl = 0;
                if (e) decide_split(e, (int)(((long)net->B * b.ih * b.iw + stride - 1) / stride) * (ocp / bn), 1, sp, cl);
                if (!e || (sp > 1 && !cl)) {          // no such kernel, or a workspace split-K: the pool stays a kernel of its own
                    ipool = 0;
                    stride = umma_conv::kBM;
                    key.pool1 = false;
                    e = find_umma(key);
                }
            }
            const bool fuse1 = ipool != 0;
            if (e && !fuse && !b.out_bf16 && ocp == bn && b.oc % 4 != 0 && !kn.no_stage_out && !kn.latency) {
                UmmaKey key_so = key;            // narrow odd rows: stage the output tile in smem
                key_so.stage_out = true;
                const UmmaEntry *e_so = find_umma(key_so);
                if (e_so) e = e_so;
            }
            int mt = 1;
            if (e && !kn.no_mt2 && !fuse1) {
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
                b.pool_fused = fuse || fuse1;
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
                    p.num_m_tiles = (p.M_total + stride * mt - 1) / (stride * mt);
                    p.OH = b.ih; p.OW = b.iw;
                }
                p.m_stride = stride;
                p.leaky = b.leaky;
                p.round_tf32 = b.round_tf32;
                p.tma_store = kn.no_tma_store ? 0 : 1;
                p.pool_first = fuse && all_scales_positive(b) && !kn.no_pool_first ? 1 : 0;
                // split-K when the output tiles alone cannot occupy the SMs (small batches)
                const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
                decide_split(e, tiles_mn, mt, p.splits, p.cluster_splitk);
                p.pool_stride = ipool;
                const int tiles = tiles_mn * p.splits;       // split-K: one resident CTA per work item (tiles <= SMs)
                b.grid = tiles < net->num_sms ? tiles : net->num_sms;
                // CTA pairs (conv_pair.cuh, tcgen05.mma.cta_group::2) for the big im2col blocks: a third to a
                // quarter less L2->SMEM traffic per flop.  The main launch covers whole waves of pair tiles;
                // the remaining rows go to a tail launch of the single-CTA kernel in 128 x 128 tiles.
                b.pair = nullptr;
                b.umma_tail = nullptr;
                const int sms = net->num_sms;
                if (!fuse && ipool != 2 && p.splits == 1 && bn == 256 && swz == 128 && kbs == 1 && !e->key.stage_out && ocp == b.oc &&
                    sms % 2 == 0 && !kn.no_pair) {
                    // measured (batch 512, whole step): MT = 1 everywhere 84.4 k img/s, MT = 2 for K >= 4096 83.7 k,
                    // MT = 2 everywhere 82.8 k, no pairs 82.5 k -- two accumulator stages and six smem stages are
                    // worth more than the last quarter of traffic
                    const int pmt = fuse1 ? 1 : kn.pair_mt;
                    const PairEntry *pe = find_pair(in_bf16, b.out_bf16, pmt, 256, fuse1);
                    UmmaKey kt{in_bf16, b.out_bf16, 128, swz, kbs, true, false, 1, false, fuse1};
                    const UmmaEntry *et = find_umma(kt);
                    const int pairs = sms / 2;
                    const int rows_per = 2 * pmt * stride;
                    // A tail launch costs a launch gap plus one whole K loop however few tiles it has.  Behind the long K loops of
                    // the 13x13x512/1024 blocks (a wave of pair tiles takes 100-200 us) that is cheaper than a mostly empty
                    // last wave; behind a short one (block 5: 72 K blocks, 25 us per wave; its two tails measured 78 us for
                    // 2.4 % of the rows) it is not: there the pairs take everything, the last tile ragged.
                    const bool short_k = ((long)b.kh * b.kw * b.ic < 4096 && !kn.pair_force_tail) || kn.pair_no_tail;
                    const int m_all = short_k ? (p.M_total + rows_per - 1) / rows_per : p.M_total / rows_per;      // pair tiles (whole ones when tails follow)
                    const int tiles_p = m_all * p.num_n_tiles;
                    const int min_tiles = kn.pair_min_tiles >= 0 ? kn.pair_min_tiles : 2 * pairs;
                    if (pe && et && tiles_p >= min_tiles && tiles_p > 0) {
                        // stop at the last full wave when the partial one would be less than half full
                        int main_m = m_all;
                        if (!short_k && tiles_p > pairs && tiles_p % pairs != 0 && (tiles_p % pairs) * 2 < pairs) main_m = (tiles_p / pairs) * pairs / p.num_n_tiles;
                        b.pair = pe;
                        pairconv::Params &q = b.pp;
                        q = pairconv::Params{};
                        q.H = b.ih; q.W = b.iw; q.B = net->B; q.OC = b.oc;
                        q.KH = b.kh; q.KW = b.kw; q.PT = b.pt; q.PL = b.pl;
                        q.c_chunks = c_chunks;
                        q.num_m_tiles = main_m; q.num_n_tiles = p.num_n_tiles;
                        q.M_total = p.M_total; q.m_base = 0; q.m_stride = stride;
                        q.leaky = b.leaky; q.round_tf32 = b.round_tf32;
                        const int t = main_m * p.num_n_tiles;
                        b.grid = 2 * (t < pairs ? t : pairs);
                        int rows_done = main_m * rows_per < p.M_total ? main_m * rows_per : p.M_total;
                        // whole pair tiles past the full waves: pairs again, in 256 x 128 tiles so that they
                        // spread over twice as many pairs (block 7: 20 tiles -> 40 of the 74 pairs, 100 -> 50 us)
                        b.pair_tail = nullptr;
                        const PairEntry *pt = pmt == 1 ? find_pair(in_bf16, b.out_bf16, 1, 128, fuse1) : nullptr;
                        const int m_pt = (p.M_total - rows_done) / (2 * stride);
                        if (pt && m_pt > 0 && !kn.no_pair_tail) {
                            b.pair_tail = pt;
                            b.pp_tail = q;
                            b.pp_tail.m_base = rows_done;
                            b.pp_tail.num_m_tiles = m_pt;
                            b.pp_tail.num_n_tiles = ocp / 128;
                            const int tt = m_pt * (ocp / 128);
                            b.grid_pair_tail = 2 * (tt < pairs ? tt : pairs);
                            rows_done += m_pt * 2 * stride;
                        }
                        const int rows_left = p.M_total - rows_done;
                        if (rows_left > 0) {
                            const int m_left = (rows_left + stride - 1) / stride;
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
                if (!b.pair && !fuse && !fuse1 && mt == 2 && p.splits == 1 && !e->key.stage_out && tiles > 2 * sms && tiles % sms != 0 &&
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

}  // namespace
