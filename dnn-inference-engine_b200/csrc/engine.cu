// engine.cu -- libdnn_b200.so: the C ABI declared in include/dnn_b200.h.
//   * op-level entry points with the reference's signatures (proj3/dnn_openblas.c)
//   * the resident network: weight prepack, HBM-resident activations, the forward launch sequence
//     (direct, programmatically chained, or replayed from a CUDA graph), host-buffer pipelining,
//     preprocessing and detection decoding around it.
// Host-side only.  Planning lives in planner.cuh, the kernel dispatch tables in launchers.cuh, the
// kernels in conv_umma.cuh / conv_pair.cuh / conv_rowpair.cuh / conv0_umma.cuh (tcgen05, TMA),
// ops_simt.cuh and postproc.cuh.
#include <cmath>
#include <cstdarg>
#include <algorithm>
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

#include "host_util.cuh"
#include "launchers.cuh"
#include "planner.cuh"

namespace {

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
        p.scale = b.d_scale; p.shift = b.d_shift; p.prepool = prepool; p.out_elems = b.conv_out_elems;
        const bool u8 = net->cur_in_u8 && &b == &net->blocks.front();
        if (u8) { p.lut = net->d_lut_conv0; p.swap_rb = net->cur_swap_rb; }
        CK(launch_conv0(p, b.op_bf16, b.out_bf16, b.grid, st, u8));
    } else if (b.family == DNNB_KERNEL_UMMA_ROWPAIR) {
        if (b.map_a_base != in) {
            if (encode_map_a(net, b, in)) return 1;
        }
        rowpair::Params p = b.rp;
        p.scale = b.d_scale; p.shift = b.d_shift; p.out = conv_out; p.prepool = prepool; p.out_elems = b.conv_out_elems;
        CK((b.rp_linear ? b.rowpair->launch_linear : b.rowpair->launch)(b.map_a, b.map_b, p, b.lin_maps, b.grid, st));
    } else {
        if (b.map_a_base != in) {
            if (encode_map_a(net, b, in)) return 1;
        }
        umma_conv::Params p = b.up;
        p.scale = b.d_scale; p.shift = b.d_shift; p.out = conv_out; p.prepool = prepool; p.out_elems = b.conv_out_elems;
        p.workspace = b.d_workspace; p.counters = b.d_counters; p.timeline = b.d_timeline;
        if (b.pair) {
            pairconv::Params q = b.pp;
            q.scale = b.d_scale; q.shift = b.d_shift; q.out = conv_out; q.prepool = prepool; q.out_elems = b.conv_out_elems;
            if (q.num_m_tiles > 0) CK(b.pair->launch(b.map_a, b.map_b_pair, q, b.grid, st));
            if (b.pair_tail) {
                pairconv::Params t = b.pp_tail;
                t.scale = b.d_scale; t.shift = b.d_shift; t.out = conv_out; t.prepool = prepool; t.out_elems = b.conv_out_elems;
                CK(b.pair_tail->launch(b.map_a, b.map_b_pair_tail, t, b.grid_pair_tail, st));
            }
        } else {
            CK(b.umma->launch(b.map_a, b.map_b, p, p.splits > 1 ? -b.grid : b.grid, st));
        }
        if (b.umma_tail) {
            umma_conv::Params q = b.up_tail;
            q.scale = b.d_scale; q.shift = b.d_shift; q.out = conv_out; q.prepool = prepool; q.out_elems = b.conv_out_elems;
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

// The device-entry calls (dnnb_net_forward_device with dev_out == NULL, forward_u8_device, detect_device) use
// slot 0 of the host pipeline's double buffers: d_in2[0] as fp32 staging and the last block's own buffer
// (= d_outf[0]) as the result.  They order themselves after the H2D / D2H copies the *_async calls still have
// in flight on those buffers, and publish their own completion to the pipeline the same way.
bool stream_is_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return true; }
    return cs != cudaStreamCaptureStatusNone;
}
// (a caller capturing the call into its own graph orders it itself: events from outside a capture cannot be waited on)
int slot0_acquire(dnnb_net *net, cudaStream_t st, bool uses_in, bool uses_out) {
    if ((!uses_in && !uses_out) || stream_is_capturing(st)) return 0;
    if (uses_in) CK(cudaStreamWaitEvent(st, net->in_free[0], 0));
    if (uses_in) CK(cudaStreamWaitEvent(st, net->in_ready[0], 0));
    if (uses_out) CK(cudaStreamWaitEvent(st, net->out_free[0], 0));
    return 0;
}
int slot0_release(dnnb_net *net, cudaStream_t st, bool uses_in, bool uses_out) {
    if ((!uses_in && !uses_out) || stream_is_capturing(st)) return 0;
    if (uses_in) CK(cudaEventRecord(net->in_free[0], st));
    if (uses_out) CK(cudaEventRecord(net->out_free[0], st));
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
    {
        std::string err;
        if (read_knobs(net->knobs, net->knobs_active, err)) return fail("%s", err.c_str());
        if (net->latency_override >= 0) {        // dnnb_net_set_latency wins over DNNB_LATENCY_MODE
            net->knobs.latency = net->latency_override != 0;
        }
    }
    net->pdl = (net->knobs.latency || net->knobs.pdl_all) && !net->knobs.no_pdl;
    net->use_graph = net->knobs.latency && !net->knobs.no_graph;
    // ---- geometry and plan of every block (no device memory yet)
    // Latency plan: the N-tile width decides how a block is cut (tiles x K slices) and through what the slices exchange
    // their partial sums.  Every width the kernels exist for is planned and the modelled time of the plan compared
    // (latency_cost); a width given by dnnb_net_set_latency_widths or DNNB_LAT_BN is taken as it is.  `crowding`:
    // the block is followed by a cluster split-K kernel, whose CTAs can start (prologue, first weight tiles) while
    // this one still runs only on SMs it leaves free -- and a cluster needs a whole group of them.
    auto choose_latency_width = [&](size_t i, bool blk_in_bf16, bool crowding) {
        Block &b = net->blocks[i];
        const int forced_bn = i < net->lat_widths.size() && net->lat_widths[i] > 0 ? net->lat_widths[i] : net->knobs.lat_bn;
        if (forced_bn > 0) {
            b.lat_bn = forced_bn;
            plan_block(net, i, blk_in_bf16);
            if (b.family == DNNB_KERNEL_SIMT) {          // no tcgen05 kernel of that width: the planner's own width
                b.lat_bn = 0;
                plan_block(net, i, blk_in_bf16);
            }
            return;
        }
        auto cost_of = [&]() {
            double c = latency_cost(b, net->num_sms);
            if (c >= 0.0 && crowding) {
                const int ctas = b.up.num_m_tiles * b.up.num_n_tiles * (b.up.splits > 1 ? b.up.splits : 1);
                c += 2.0 * (ctas < net->num_sms ? (double)ctas / net->num_sms : 1.0);
            }
            return c;
        };
        int best_bn = 0;
        double best = 0.0;
        const int cands[4] = {0, 128, 64, 32};       // 0: the planner's own width (up to 256)
        for (int c = 0; c < 4; ++c) {
            b.lat_bn = cands[c];
            plan_block(net, i, blk_in_bf16);
            if (c == 0) {
                best = cost_of();
                if (best < 0.0) break;               // not a block of the generic kernel: nothing to choose
                continue;
            }
            if (b.bn != cands[c]) continue;          // the cap did not bind: same plan as a wider candidate
            const double cost = cost_of();
            if (cost < 0.0) continue;
            if (cost + 0.2 <= best) {                // a tie goes to the wider tile
                best = cost;
                best_bn = cands[c];
            }
        }
        b.lat_bn = best_bn;
        plan_block(net, i, blk_in_bf16);
        if (net->knobs.verbose_plan)
            fprintf(stderr, "dnnb plan: block %zu latency width %d (bn %d, splits %d, cluster %d%s) modelled %.1f us\n", i, best_bn,
                    b.bn, b.up.splits, b.up.cluster_splitk, crowding ? ", crowding its successor" : "", best);
    };
    std::vector<char> blk_in_bf16(net->blocks.size(), 0);
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
        blk_in_bf16[i] = in_bf16;
        if (i == 0 && net->precision == DNNB_BF16) {
            // conv0 reads the fp32 image itself; any other tcgen05 first block gets a bf16 cast
            plan_block(net, i, false);
            if (b.family != DNNB_KERNEL_UMMA_CONV0) {
                plan_block(net, i, true);
                if (b.family == DNNB_KERNEL_SIMT) plan_block(net, i, false);
            }
        } else if (net->knobs.latency) {
            choose_latency_width(i, in_bf16, false);
        } else {
            plan_block(net, i, in_bf16);
        }
        in_bf16 = b.out_bf16;
        h = b.oh; w = b.ow;
    }
    if (net->knobs.latency) {
        // second look at the blocks that are followed by a cluster split-K kernel (a block's plan does not depend on its
        // neighbours' plans, so one pass is enough)
        for (size_t i = 0; i + 1 < net->blocks.size(); ++i) {
            const Block &next = net->blocks[i + 1];
            const bool next_cluster = (next.family == DNNB_KERNEL_UMMA_IM2COL || next.family == DNNB_KERNEL_UMMA_TILED) &&
                                      next.up.splits > 1 && next.up.cluster_splitk;
            if (next_cluster && !(i == 0 && net->precision == DNNB_BF16)) choose_latency_width(i, blk_in_bf16[i] != 0, true);
        }
    }
    // ---- device memory, constants, prepacked weights, tensor maps
    for (size_t i = 0; i < net->blocks.size(); ++i) {
        Block &b = net->blocks[i];
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
            if (b.up.splits > 1 && !b.up.cluster_splitk) {
                const size_t ws = (size_t)b.up.splits * b.up.num_m_tiles * (b.umma->key.mt) * umma_conv::kBM * ocp * sizeof(float);
                const size_t nc = (size_t)b.up.num_m_tiles * b.up.num_n_tiles * 2 * sizeof(int);
                CK(cudaMalloc(&b.d_workspace, ws));
                CK(cudaMalloc(&b.d_counters, nc));
                CK(cudaMemset(b.d_workspace, 0, ws));
                CK(cudaMemset(b.d_counters, 0, nc));
            }
        }
        net->launches += 1 + (b.pool_separate ? 1 : 0);
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
    // TMA tensor maps, bulk copies and the vector stores of the epilogues need 16-byte aligned tensors
    if ((reinterpret_cast<uintptr_t>(dev_in) & 15) || (reinterpret_cast<uintptr_t>(dev_out) & 15))
        return fail("dev_in and dev_out must be 16-byte aligned");
    DeviceGuard g(net->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : net->stream;
    net->last_final = dev_out ? nullptr : static_cast<const float *>(net->blocks.back().d_out);
    net->last_in = dev_in;
    const bool own_out = dev_out == nullptr;
    if (own_out && slot0_acquire(net, st, false, true)) return 1;
    if (net->use_graph && net->prof_block < 0) {
        for (dnnb_net::GraphSlot &g : net->graphs)
            if (g.in == dev_in && g.out == dev_out) {
                g.used = ++net->graph_tick;
                CK(cudaGraphLaunch(g.exec, st));
                return own_out ? slot0_release(net, st, false, true) : 0;
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
                return own_out ? slot0_release(net, st, false, true) : 0;
            }
            net->use_graph = false;       // capture is not available here (legacy stream, driver): launch directly from now on
        }
        net->warm_in = dev_in;
        net->warm_out = dev_out;
    }
    if (forward_impl(net, dev_in, dev_out, st)) return 1;
    return own_out ? slot0_release(net, st, false, true) : 0;
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
    if ((reinterpret_cast<uintptr_t>(dev_u8) & 15) || (reinterpret_cast<uintptr_t>(dev_out) & 15))
        return fail("dev_u8 and dev_out must be 16-byte aligned");
    DeviceGuard g(net->device);
    if (ensure_u8(net)) return 1;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : net->stream;
    const bool fused = u8_fused(net, 0, 0);
    if (slot0_acquire(net, st, !fused, dev_out == nullptr)) return 1;
    if (!fused) CK(launch_preprocess(net, static_cast<const uint8_t *>(dev_u8), net->d_in2[0], bgr ? 1 : 0, st));
    net->last_final = dev_out ? nullptr : static_cast<const float *>(net->blocks.back().d_out);
    net->last_in = fused ? nullptr : net->d_in2[0];
    if (fused ? forward_impl(net, dev_u8, dev_out, st, true, bgr ? 1 : 0) : forward_impl(net, net->d_in2[0], dev_out, st)) return 1;
    return slot0_release(net, st, !fused, dev_out == nullptr);
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
    if (reinterpret_cast<uintptr_t>(dev_in) & 15) return fail("dev_in must be 16-byte aligned");
    DeviceGuard g(net->device);
    if (ensure_detect(net)) return 1;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : net->stream;
    const void *in = dev_in;
    bool fused = false;
    if (is_u8) {
        if (ensure_u8(net)) return 1;
        fused = u8_fused(net, 0, 0);
    }
    const bool stage = is_u8 && !fused;
    if (slot0_acquire(net, st, stage, true)) return 1;
    if (stage) {
        CK(launch_preprocess(net, static_cast<const uint8_t *>(dev_in), net->d_in2[0], bgr ? 1 : 0, st));
        in = net->d_in2[0];
    }
    net->last_final = static_cast<const float *>(net->blocks.back().d_out);
    net->last_in = fused ? nullptr : in;
    if (forward_impl(net, in, nullptr, st, fused, bgr ? 1 : 0)) return 1;
    CK(launch_detect(net, net->last_final, dev_dets, dev_counts, st));
    return slot0_release(net, st, stage, true);
}

int dnnb_net_sync(dnnb_net *net) {
    if (check_net(net, false)) return 1;
    DeviceGuard g(net->device);
    CK(cudaStreamSynchronize(net->stream));
    if (net->copy_in) CK(cudaStreamSynchronize(net->copy_in));
    if (net->copy_out) CK(cudaStreamSynchronize(net->copy_out));
    return 0;
}

int dnnb_net_next_slot(dnnb_net *net, int *slot) {
    if (check_net(net, true) || !slot) return fail("bad argument");
    *slot = net->flip;
    return 0;
}

int dnnb_net_wait_slot(dnnb_net *net, int slot) {
    if (check_net(net, true)) return 1;
    if (slot < 0 || slot > 1) return fail("slot must be 0 or 1");
    DeviceGuard g(net->device);
    CK(cudaEventSynchronize(net->out_free[slot]));      // recorded after the slot's D2H copies
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
    info[12] = b.family == DNNB_KERNEL_UMMA_CONV0 ? b.c0.num_tiles
               : b.family == DNNB_KERNEL_UMMA_ROWPAIR ? b.rp.num_tiles
               : b.pair ? b.pp.num_m_tiles * b.pp.num_n_tiles
               : b.umma ? b.up.num_m_tiles * b.up.num_n_tiles * b.up.splits : 0;
    info[13] = b.umma && !b.pair ? b.up.splits : 1;
    info[14] = b.umma && !b.pair && b.up.splits > 1 && b.up.cluster_splitk ? 1 : 0;
    info[15] = 0;
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

int dnnb_net_time_forward(dnnb_net *net, int warmup, int iters, float *median_us) {
    if (check_net(net, true)) return 1;
    if (!median_us || iters <= 0 || warmup < 0) return fail("bad argument");
    DeviceGuard g(net->device);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<float> us;
    us.reserve(iters);
    for (int it = 0; it < warmup + iters; ++it) {
        CK(cudaEventRecord(e0, net->stream));
        if (dnnb_net_forward_device(net, net->d_in, nullptr, nullptr)) return 1;      // graph replay under the latency plan
        CK(cudaEventRecord(e1, net->stream));
        CK(cudaEventSynchronize(e1));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, e0, e1));
        if (it >= warmup) us.push_back(t * 1e3f);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    std::sort(us.begin(), us.end());
    *median_us = us[us.size() / 2];
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

int dnnb_net_set_latency(dnnb_net *net, int mode) {
    if (check_net(net, false)) return 1;
    if (net->finalized) return fail("dnnb_net_set_latency must be called before dnnb_net_finalize");
    if (mode < -1 || mode > 1) return fail("mode must be -1 (environment decides), 0 or 1");
    net->latency_override = mode;
    return 0;
}

int dnnb_net_set_latency_widths(dnnb_net *net, const int *widths, int n) {
    if (check_net(net, false)) return 1;
    if (net->finalized) return fail("dnnb_net_set_latency_widths must be called before dnnb_net_finalize");
    if (n < 0 || (n > 0 && !widths)) return fail("bad argument");
    for (int i = 0; i < n; ++i)
        if (widths[i] != 0 && widths[i] != 32 && widths[i] != 64 && widths[i] != 128 && widths[i] != 256)
            return fail("width %d of block %d: 0 (modelled choice), 32, 64, 128 or 256", widths[i], i);
    net->lat_widths.assign(widths, widths + n);
    return 0;
}

int dnnb_net_knobs(dnnb_net *net, char *buf, int buf_len) {
    if (check_net(net, true) || !buf || buf_len < 1) return fail("bad argument");
    std::string text = net->knobs_active;
    if (net->latency_override >= 0) {
        if (!text.empty()) text += ",";
        text += std::string("set_latency=") + (net->latency_override ? "1" : "0");
    }
    bool any_width = false;
    for (int w : net->lat_widths) any_width = any_width || w != 0;
    if (any_width) {
        if (!text.empty()) text += ",";
        text += "set_latency_widths=";
        for (size_t i = 0; i < net->lat_widths.size(); ++i) text += (i ? "/" : "") + std::to_string(net->lat_widths[i]);
    }
    if ((int)text.size() + 1 > buf_len) return fail("buffer too small: %zu bytes needed", text.size() + 1);
    memcpy(buf, text.c_str(), text.size() + 1);
    return 0;
}

int dnnb_net_debug_timeline(dnnb_net *net, int block, unsigned long long *host_out, int *ctas) {
    if (check_net(net, true)) return 1;
    if (block < 0 || block >= (int)net->blocks.size() || !ctas) return fail("bad block index or argument");
    Block &b = net->blocks[block];
    if (!b.umma || b.pair) return fail("block %d does not run the generic tcgen05 kernel", block);
    DeviceGuard g(net->device);
    CK(dnnb_sync_streams(net));
    const size_t bytes = (size_t)b.grid * DNNB_TIMELINE_SLOTS * sizeof(unsigned long long);
    if (!host_out) {                         // arm: allocate + clear; graphs recorded without the pointer are dropped
        if (!b.d_timeline) CK(cudaMalloc(&b.d_timeline, bytes));
        CK(cudaMemset(b.d_timeline, 0, bytes));
        for (dnnb_net::GraphSlot &gs : net->graphs) cudaGraphExecDestroy(gs.exec);
        net->graphs.clear();
        net->warm_in = nullptr;
        net->warm_out = nullptr;
        *ctas = b.grid;
        return 0;
    }
    if (!b.d_timeline) return fail("timeline of block %d was never armed", block);
    CK(cudaMemcpy(host_out, b.d_timeline, bytes, cudaMemcpyDeviceToHost));
    *ctas = b.grid;
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
        if (b.d_timeline) cudaFree(b.d_timeline);
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
    static_assert(DNNB_TIMELINE_SLOTS == umma_conv::kTimelineSlots, "timeline slots");
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
