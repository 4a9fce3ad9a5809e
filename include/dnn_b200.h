/* dnn_b200.h -- C ABI of libdnn_b200.so, the sm_100a replacement for the native op
 * libraries of tink-expo/dnn-inference-engine proj3 (libdnn_openblas.so / libdnn_avx.so /
 * libdnn_cublas.so / libdnn_cuda.so, proj3/Makefile:1-15).
 *
 * Two groups of entry points:
 *
 *  1. Op-level functions with the names and argument lists of proj3/dnn_openblas.c, taking
 *     HOST pointers.  proj3/dnn_openblas.py binds them unchanged
 *     (ctypes.cdll.LoadLibrary, proj3/dnn_openblas.py:9); each call uploads, runs one CUDA
 *     kernel and downloads.  They exist for op-by-op parity and for the debug
 *     (layer_N.npy) mode of the host engine; they are not the fast path.
 *
 *  2. A resident network API (dnnb_net_*): weights are uploaded and prepacked once,
 *     activations stay in HBM across all conv blocks, and one forward call runs the whole
 *     fused chain.  This is what dnn_b200.DnnInferenceEngine.run uses.
 *
 * Conventions: every function returns 0 on success, non-zero on failure, and
 * dnnb_last_error() then describes the failure (the reference has no error convention at all:
 * void returns, unchecked CUDA statuses -- SURVEY.md 8b).  All tensors are NHWC float32 on
 * the host side.  Host buffers stay owned by the caller and are not retained after a call.
 * Inputs are never written (the reference's batch_norm overwrites its variance argument,
 * proj3/dnn_openblas.c:48-50; this library does not).  There is no CPU fallback: without a
 * CUDA device every compute entry point fails.
 */
#ifndef DNN_B200_H
#define DNN_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dnnb_net dnnb_net; /* opaque network handle */

/* arithmetic the conv blocks run in */
#define DNNB_TF32 0 /* tcgen05 kind::tf32, fp32 activations rounded to TF32 (RN) between blocks */
#define DNNB_BF16 1 /* tcgen05 kind::f16 (bf16 operands), bf16 activations between blocks        */
#define DNNB_FP32 2 /* CUDA-core fp32 kernels only (debug / tight-parity path)                    */

/* kernel family chosen for a block, reported by dnnb_net_block_info */
#define DNNB_KERNEL_SIMT 0      /* CUDA-core direct convolution                 */
#define DNNB_KERNEL_UMMA_TILED 1 /* tcgen05 implicit GEMM, TMA tiled boxes, fused 2x2/2 pool */
#define DNNB_KERNEL_UMMA_IM2COL 2 /* tcgen05 implicit GEMM, TMA im2col loads     */
#define DNNB_KERNEL_UMMA_ROWPAIR 4 /* tcgen05, patch-resident row-pair tiles, resident weights, fused pool */
#define DNNB_KERNEL_UMMA_PAIR 5 /* tcgen05.mma.cta_group::2: CTA pairs share each weight tile (im2col blocks, big batches) */
#define DNNB_KERNEL_UMMA_CONV0 3 /* tcgen05, 3-channel first layer from the raw image, patch-resident, fused pool */

const char *dnnb_last_error(void);
int dnnb_version(void);
int dnnb_device_count(int *count);

/* ---- 1. op-level, reference signatures (host pointers) --------------------------------- */

/* proj3/dnn_openblas.c:160-194.  in_layer is the zero-padded input [batch,ih,iw,ic]; col is
 * the reference's im2col scratch and is ignored; kernel_r is [ic*kh*kw, od] with
 * k = c*kh*kw + di*kw + dj (proj3/dnn_openblas.py:166-167).  Unlike the reference (:170)
 * batch > 1 is handled correctly. */
int conv2d_mul(float *in_layer, float *col, float *kernel_r, float *result,
               int batch, int oh, int ow, int od, int ih, int iw, int ic,
               int kh, int kw, int sh, int sw);
/* proj3/dnn_openblas.c:9-38 */
int bias_add(float *in_layer, float *biases, float *result,
             int shape_b, int shape_h, int shape_w, int shape_d);
/* proj3/dnn_openblas.c:40-65; variance is NOT modified; batch > 1 handled */
int batch_norm(float *in_layer, float *mean, float *variance, float *gamma, float epsilon,
               float *result, int batch, int oh, int ow, int od);
/* proj3/dnn_openblas.c:196-234; in_layer is padded with finfo(float32).min by the caller
 * (proj3/dnn_openblas.py:232-235).  The reference's host passes four trailing pad ints that
 * its C ignores (proj3/dnn_openblas.py:242); they are accepted and ignored here too. */
int max_pool2d(float *in_layer, float *result,
               int batch, int oh, int ow, int od, int ih, int iw, int ic,
               int kh, int kw, int sh, int sw);
/* proj3/dnn_openblas.c:236-254 */
int leaky_relu(float *in_layer, float *result, int batch, int oh, int ow, int od);

/* ---- 2. resident network API ----------------------------------------------------------- */

/* Creates an empty network on `device` for inputs [batch,in_h,in_w,in_c] float32 NHWC. */
int dnnb_net_create(int device, int precision, int batch, int in_h, int in_w, int in_c,
                    dnnb_net **out);

/* Appends one fused block  Conv2D -> [+bias] -> [BatchNorm] -> [LeakyReLU(0.1)] -> [MaxPool2D],
 * i.e. the node run of proj3/yolov2tiny.py:30-34.  kernel is HWIO [kh,kw,ic,oc]; stride 1.
 * biases may be NULL; mean/variance/gamma NULL => no batch norm; pool_k == 0 => no pool.
 * padding_same / pool_same: 1 = 'SAME', 0 = 'VALID' (proj3/dnn_openblas.py:127-142).
 * All arrays are copied; none is modified. */
int dnnb_net_add_block(dnnb_net *net, const float *kernel, int kh, int kw, int ic, int oc,
                       int padding_same, const float *biases,
                       const float *mean, const float *variance, const float *gamma, float epsilon,
                       int leaky, int pool_k, int pool_s, int pool_same);

/* Chooses a kernel per block, prepacks weights, allocates the HBM-resident activations.
 * Environment, read here: DNNB_LATENCY_MODE=1 plans for small-batch latency -- split-K over co-resident
 * CTAs (deterministic, but a different fp32 summation order than the default plan), kernels chained by
 * programmatic dependent launch, and dnnb_net_forward_device replayed from a CUDA graph per
 * (dev_in, dev_out) pair from the second call on (the first call launches directly and sets function
 * attributes and tensor maps).  DNNB_NO_PDL / DNNB_NO_GRAPH switch the last two off.
 * Residency: blocks whose K slices are reduced through the global workspace (the 13x13x1024 blocks at batch 1) wait
 * for their sibling CTAs inside the kernel; the launch is cooperative, or -- under programmatic dependent launch, where
 * that attribute is not available -- relies on its grid (<= one CTA per SM) having the device to itself: run ONE
 * latency-plan network per device at a time (a sibling that never becomes resident ends in a bounded wait and a trap
 * that the next call reports, not in a hang).  Blocks split over thread-block clusters carry no such assumption.
 * Streams: the *_async host calls pipeline on three streams owned by the network.  The device-entry calls run on
 * the caller's stream (or the network's own when NULL); those of them that use the network's own staging /
 * result buffers (dev_out == NULL, uint8 input that needs a preprocessing pass, dnnb_net_detect_device) order
 * themselves after the host pipeline's copies on those buffers and are in turn waited for by later *_async calls. */
int dnnb_net_finalize(dnnb_net *net);

/* Selects the small-batch latency plan for this network without touching the environment: 1 = on, 0 = off,
 * -1 = DNNB_LATENCY_MODE decides (the default).  Call before dnnb_net_finalize. */
int dnnb_net_set_latency(dnnb_net *net, int mode);

/* Latency plan only: the N-tile width (32, 64, 128, 256; 0 = the planner's modelled choice, the default) of the first
 * n blocks.  The width decides how a block is cut into output tiles and K slices and how the slices exchange their
 * partial sums; the planner compares modelled times (planner.cuh, latency_cost), tools/tune_latency.py measures the
 * alternatives on the device.  A width the kernels do not exist for falls back to the planner's own.  Results of
 * different widths differ by fp32 summation order where the number of K slices differs.  Call before
 * dnnb_net_finalize; dnnb_net_knobs reports a non-default choice as set_latency_widths=a/b/... */
int dnnb_net_set_latency_widths(dnnb_net *net, const int *widths, int n);

/* Environment switches.  dnnb_net_finalize reads the DNNB_* variables (DESIGN.md section 4c lists them; apart
 * from DNNB_LATENCY_MODE they force a particular kernel for tests and ablation measurements) and FAILS on a
 * DNNB_* name it does not know or a value it cannot parse, so that a stale variable in a shell cannot silently
 * change the plan.  dnnb_net_knobs writes the switches that are not at their default as "NAME=value,..." (an
 * empty string for the default plan) into buf. */
int dnnb_net_knobs(dnnb_net *net, char *buf, int buf_len);

int dnnb_net_num_blocks(dnnb_net *net, int *n);
int dnnb_net_out_shape(dnnb_net *net, int *b, int *h, int *w, int *c);

/* Host in / host out: H2D copy, all blocks, D2H copy; returns when host_out is valid. */
int dnnb_net_forward(dnnb_net *net, const float *host_in, float *host_out);

/* Device in / device out (float32 NHWC in HBM), asynchronous on `stream` (a cudaStream_t, or
 * NULL for the network's own stream).  Both pointers must be 16-byte aligned.  dev_out may be NULL: the result then stays in the
 * network's last activation buffer (see dnnb_net_block_output_ptr). */
int dnnb_net_forward_device(dnnb_net *net, const void *dev_in, void *dev_out, void *stream);

/* As dnnb_net_forward but only enqueues (H2D, blocks, D2H) on the network's own stream and
 * returns; host_in / host_out should be page-locked (dnnb_host_alloc) and must stay valid until
 * dnnb_net_sync returns.  Lets one host thread keep several devices busy. */
int dnnb_net_forward_async(dnnb_net *net, const float *host_in, float *host_out);

/* uint8 frames: the device half of proj3/__init__.py:8-12 (resize_input) for frames that already are
 * in_h x in_w -- x = float32(u8 / 255.), and with bgr != 0 the channel reversal imsz[:, :, ::-1] --
 * runs as one extra kernel in front of block 0, so only 3 bytes per pixel cross PCIe instead of 12.
 * Bit-identical to feeding dnnb_net_forward the numpy-preprocessed float32 array.  3-channel inputs only.
 * The async form pipelines like dnnb_net_forward_async (page-locked buffers, dnnb_net_sync). */
int dnnb_net_forward_u8_async(dnnb_net *net, const unsigned char *host_u8, int bgr, float *host_out);
int dnnb_net_forward_u8_device(dnnb_net *net, const void *dev_u8, int bgr, void *dev_out, void *stream);

/* Frames of ANY size: the whole of proj3/__init__.py:8-12 (resize_input) on the device --
 * cv2.resize(frame, (in_w, in_h)) [8-bit INTER_LINEAR, bit-identical to OpenCV's fixed-point
 * path], / 255., BGR -> RGB, float32 -- as the one kernel in front of block 0.  host_frames is
 * uint8 [batch][src_h][src_w][3].  dnnb_net_detect_frames_async (section 3) also decodes the boxes. */
int dnnb_net_forward_frames_async(dnnb_net *net, const unsigned char *host_frames, int src_h, int src_w,
                                  int bgr, float *host_out);
/* op-level form, host in / host out: float32 [batch][dst_h][dst_w][3] */
int dnnb_resize_input(int device, const unsigned char *frames_host, int batch, int src_h, int src_w,
                      int dst_h, int dst_w, int bgr, float *out_host);

/* Waits for everything queued on the network's own streams. */
int dnnb_net_sync(dnnb_net *net);

/* Streaming: the *_async host calls alternate between two pipeline slots (device-side input / output buffers).
 * dnnb_net_next_slot tells which slot the NEXT *_async call will use; dnnb_net_wait_slot returns once the
 * results of the most recent job of `slot` are in the caller's host buffers -- so a loop can wait for job k - 1
 * while job k runs (streaming.StreamingDetector; the frame loop of proj1/__init__.py:71-137). */
int dnnb_net_next_slot(dnnb_net *net, int *slot);
int dnnb_net_wait_slot(dnnb_net *net, int slot);

/* Debug / layer-wise comparison: copies block `block`'s output (after its pool, if any) to
 * host as float32 [b,h,w,c]; shape via dnnb_net_block_info. */
int dnnb_net_block_output(dnnb_net *net, int block, float *host_out);

/* Debug, operator-local pool check: re-runs block `block` (which must have its 2x2/2 pool fused in
 * a tcgen05 epilogue) on the input it saw in the last forward, this time also dumping the
 * activations the fused pool reduces -- [b, conv_h, conv_w, c] float32, after scale/shift,
 * LeakyReLU and storage rounding -- which otherwise never leave the SM.  The block's pooled
 * output is rewritten with identical values. */
int dnnb_net_debug_prepool(dnnb_net *net, int block, float *host_out);

/* info[0..15] = {kernel family, out_h, out_w, out_c, pooled(0/1: pool fused in the conv
 * epilogue), BN tile, swizzle bytes, stages, grid, separate-pool-kernel(0/1), elem bytes of the
 * stored activation, launches per forward, work items (tiles) of the main launch, K slices per
 * output tile (split-K, latency plan; 1 = none), 1 if those slices are the CTAs of one thread-block
 * cluster reduced through distributed shared memory, reserved} */
int dnnb_net_block_info(dnnb_net *net, int block, int *info);

/* Times each block's kernels with CUDA events on the network's stream: `iters` timed
 * repetitions after `warmup`; ms[block] = mean milliseconds per forward of that block. */
int dnnb_net_time_blocks(dnnb_net *net, int warmup, int iters, float *ms);

/* In-situ timing of one block inside real forwards: after dnnb_net_profile_block(net, b) every
 * forward records a CUDA-event pair around block b's kernels on the stream it runs on;
 * dnnb_net_profile_read waits for them, returns their summed duration and count, and resets.
 * block < 0 switches it off. */
int dnnb_net_profile_block(dnnb_net *net, int block);
int dnnb_net_profile_read(dnnb_net *net, float *total_ms, int *count);

/* Number of kernel launches one forward issues (for bench.py's gpu_launches). */
int dnnb_net_launches_per_forward(dnnb_net *net, int *n);

/* Debug / measurement: forwards stop after the first `blocks` blocks (<= 0: all of them).  Timing
 * forwards at increasing limits gives each block's in-situ share of the batch-1 latency, overlap
 * between programmatically chained kernels included (tools/batch1_probe.py). */
int dnnb_net_debug_block_limit(dnnb_net *net, int blocks);

/* Measurement: median device time (us, CUDA events on the network's own stream, one forward per event pair) of
 * `iters` forwards from the network's own input buffer (whatever it holds) into its own output, after `warmup`
 * untimed ones.  Under the latency plan these are the graph replays a caller gets.  What tune_latency_widths of the
 * Python binding compares candidate plans with. */
int dnnb_net_time_forward(dnnb_net *net, int warmup, int iters, float *median_us);

/* Debug / measurement: per-CTA phase stamps (%globaltimer, ns) of a block that runs the generic tcgen05 kernel.
 * host_out == NULL arms the block (later forwards record DNNB_TIMELINE_SLOTS stamps per CTA: 0 entry, 1 prologue done,
 * 2 predecessor complete, 3 first operands landed, 4 last MMA issued, 5 accumulator complete, 6 split-K siblings seen,
 * 7 result stored, and for split-K blocks 8 own partial tile written to the workspace, 9 siblings' partial sums
 * read; a phase a CTA does not go through stays 0); host_out != NULL copies [ctas][DNNB_TIMELINE_SLOTS] stamps back.
 * *ctas receives the block's grid size either way.  Where the micro-seconds of the batch-1 chain go
 * (tools/batch1_timeline.py). */
#define DNNB_TIMELINE_SLOTS 10
int dnnb_net_debug_timeline(dnnb_net *net, int block, unsigned long long *host_out, int *ctas);

int dnnb_net_destroy(dnnb_net *net);

/* ------------------------------------------------------------------------------------------
 * 3. Detection decoding on the device: the step that follows the layer stack in the reference,
 *    `postprocessing` (proj3/yolov2tiny.py:94-180) with its `iou` (:182-200),
 *    `non_maximal_suppression` (:202-226), `sigmoid` (:228) and `softmax` (:231).  Input is the
 *    [batch,13,13,125] head tensor; output per image: the label boxes that survive the score
 *    threshold (reference 0.3, compared in float32) and the greedy class-agnostic NMS (IoU
 *    threshold 0.3, compared in double), in the reference's order (score descending, stable).
 *    Class names / colours are a host-side table lookup on `cls` (dnn_b200.py).  Box edges
 *    saturate at +-2^29 (the reference has unbounded Python ints); a zero union, on which the
 *    reference raises ZeroDivisionError, suppresses the box.
 * ------------------------------------------------------------------------------------------ */
typedef struct dnnb_det {
    int cls;                      /* 0..19, Pascal VOC order of proj3/yolov2tiny.py:103 */
    int left, top, right, bottom; /* int(center -/+ roi/2), pixels of the 416x416 input */
    float score;                  /* sigmoid(tc) * max softmax */
} dnnb_det;
#define DNNB_CANDIDATES 845 /* 13 x 13 cells x 5 anchor boxes */

/* dets_host [batch][max_det], counts_host [batch] = boxes kept (only min(count, max_det) stored) */
int dnnb_postprocess(int device, const float *pred_host, int batch, float score_threshold,
                     double iou_threshold, int max_det, dnnb_det *dets_host, int *counts_host);
/* parity hooks: all 845 decoded candidates per image in (row, col, box) order, before the
 * threshold; and the threshold + sort + NMS half alone on caller-supplied candidates */
int dnnb_postprocess_candidates(int device, const float *pred_host, int batch, dnnb_det *cand_host);
int dnnb_nms(int device, const dnnb_det *cand_host, int batch, float score_threshold,
             double iou_threshold, int max_det, dnnb_det *dets_host, int *counts_host);

/* Network form: forward + decoding enqueued together; only boxes and counts cross PCIe.
 * Thresholds default to the reference's 0.3 / 0.3 and max_det to 32 per image.  The async forms
 * pipeline like dnnb_net_forward_async (page-locked buffers, dnnb_net_sync). */
int dnnb_net_set_detect(dnnb_net *net, float score_threshold, double iou_threshold, int max_det);
int dnnb_net_detect_async(dnnb_net *net, const float *host_in, dnnb_det *dets_host, int *counts_host);
int dnnb_net_detect_u8_async(dnnb_net *net, const unsigned char *host_u8, int bgr, dnnb_det *dets_host,
                             int *counts_host);
int dnnb_net_detect_frames_async(dnnb_net *net, const unsigned char *host_frames, int src_h, int src_w,
                                 int bgr, dnnb_det *dets_host, int *counts_host);
/* input (fp32 NHWC, or uint8 frames with is_u8) and outputs all in device memory */
int dnnb_net_detect_device(dnnb_net *net, const void *dev_in, int is_u8, int bgr, void *dev_dets,
                           void *dev_counts, void *stream);

/* Page-locked host memory for the end-to-end path. */
int dnnb_host_alloc(size_t bytes, void **out);
int dnnb_host_free(void *p);

/* Measures dense tcgen05 throughput (TFLOP/s) of kind::tf32 (precision 0) or kind::f16 bf16
 * (precision 1) with a register-free MMA issue loop on every SM: the tensor-pipe roofline
 * denominator for TF32, which MEASURED_PEAKS.json does not hold. */
int dnnb_measure_mma_peak(int device, int precision, float *tflops);
/* The same loop on pseudo-random operands (the burst probe multiplies zeros, which draws little
 * power) run back to back for `seconds` (<= 30): the rate the part settles at under its power cap,
 * the denominator for a kernel timed inside a long step. */
int dnnb_measure_mma_peak_sustained(int device, int precision, float seconds, float *tflops);

#ifdef __cplusplus
}
#endif
#endif
