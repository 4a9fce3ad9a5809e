"""ctypes binding of libdnn_b200.so (include/dnn_b200.h).  The library is loaded by absolute
path from next to this file (the reference loads './libdnn_X.so' relative to the cwd,
proj3/dnn_openblas.py:9).  There is no fallback: a missing library is an ImportError, and a
failing call raises RuntimeError with the library's message."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LIBDNN_B200_PATH") or os.path.join(HERE, "libdnn_b200.so")     # override: A/B runs of two builds

TF32, BF16, FP32 = 0, 1, 2
PRECISIONS = {"tf32": TF32, "bf16": BF16, "fp32": FP32}
KERNEL_NAMES = {0: "simt", 1: "umma_tiled_pool", 2: "umma_im2col", 3: "umma_conv0", 4: "umma_rowpair", 5: "umma_pair"}

if not os.path.exists(LIB_PATH):
    raise ImportError(LIB_PATH + " is missing: build it with `python dnn-inference-engine_b200/build.py` "
                      "(nvcc, sm_100a); this engine has no CPU fallback")

lib = ctypes.CDLL(LIB_PATH)
TIMELINE_SLOTS = 10            # DNNB_TIMELINE_SLOTS of include/dnn_b200.h

c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
FP = ctypes.POINTER(c_float)
IP = ctypes.POINTER(c_int)

lib.dnnb_last_error.restype = ctypes.c_char_p
lib.dnnb_last_error.argtypes = []


def _sig(name, *argtypes):
    f = getattr(lib, name)
    f.restype = c_int
    f.argtypes = list(argtypes)
    return f


_sig("dnnb_version")
_sig("dnnb_device_count", IP)
_sig("conv2d_mul", FP, FP, FP, FP, *([c_int] * 11))
_sig("bias_add", FP, FP, FP, *([c_int] * 4))
_sig("batch_norm", FP, FP, FP, FP, c_float, FP, *([c_int] * 4))
_sig("max_pool2d", FP, FP, *([c_int] * 11))
_sig("leaky_relu", FP, FP, *([c_int] * 4))
_sig("dnnb_net_create", c_int, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_void_p))
_sig("dnnb_net_add_block", c_void_p, FP, c_int, c_int, c_int, c_int, c_int, FP, FP, FP, FP, c_float,
     c_int, c_int, c_int, c_int)
_sig("dnnb_net_finalize", c_void_p)
_sig("dnnb_net_set_latency", c_void_p, c_int)
_sig("dnnb_net_set_latency_widths", c_void_p, IP, c_int)
_sig("dnnb_net_knobs", c_void_p, ctypes.c_char_p, c_int)
_sig("dnnb_net_num_blocks", c_void_p, IP)
_sig("dnnb_net_out_shape", c_void_p, IP, IP, IP, IP)
_sig("dnnb_net_forward", c_void_p, c_void_p, c_void_p)
_sig("dnnb_net_forward_async", c_void_p, c_void_p, c_void_p)
_sig("dnnb_net_forward_device", c_void_p, c_void_p, c_void_p, c_void_p)
_sig("dnnb_net_forward_u8_async", c_void_p, c_void_p, c_int, c_void_p)
_sig("dnnb_net_forward_u8_device", c_void_p, c_void_p, c_int, c_void_p, c_void_p)
_sig("dnnb_net_sync", c_void_p)
_sig("dnnb_net_next_slot", c_void_p, IP)
_sig("dnnb_net_wait_slot", c_void_p, c_int)
_sig("dnnb_net_block_output", c_void_p, c_int, FP)
_sig("dnnb_net_block_info", c_void_p, c_int, IP)
_sig("dnnb_net_debug_prepool", c_void_p, c_int, FP)
_sig("dnnb_net_time_blocks", c_void_p, c_int, c_int, FP)
_sig("dnnb_net_launches_per_forward", c_void_p, IP)
_sig("dnnb_net_debug_block_limit", c_void_p, c_int)
_sig("dnnb_net_debug_timeline", c_void_p, c_int, c_void_p, IP)
_sig("dnnb_net_time_forward", c_void_p, c_int, c_int, ctypes.POINTER(c_float))
_sig("dnnb_net_profile_block", c_void_p, c_int)
_sig("dnnb_net_profile_read", c_void_p, FP, IP)
_sig("dnnb_net_destroy", c_void_p)
_sig("dnnb_host_alloc", c_size_t, ctypes.POINTER(c_void_p))
_sig("dnnb_host_free", c_void_p)
_sig("dnnb_postprocess", c_int, c_void_p, c_int, c_float, ctypes.c_double, c_int, c_void_p, c_void_p)
_sig("dnnb_postprocess_candidates", c_int, c_void_p, c_int, c_void_p)
_sig("dnnb_nms", c_int, c_void_p, c_int, c_float, ctypes.c_double, c_int, c_void_p, c_void_p)
_sig("dnnb_net_set_detect", c_void_p, c_float, ctypes.c_double, c_int)
_sig("dnnb_net_detect_async", c_void_p, c_void_p, c_void_p, c_void_p)
_sig("dnnb_net_detect_u8_async", c_void_p, c_void_p, c_int, c_void_p, c_void_p)
_sig("dnnb_net_detect_device", c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p)
_sig("dnnb_net_forward_frames_async", c_void_p, c_void_p, c_int, c_int, c_int, c_void_p)
_sig("dnnb_net_detect_frames_async", c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p)
_sig("dnnb_resize_input", c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p)
_sig("dnnb_measure_mma_peak", c_int, c_int, FP)
_sig("dnnb_measure_mma_peak_sustained", c_int, c_int, c_float, FP)

EXPORTS = ["dnnb_last_error", "dnnb_version", "dnnb_device_count", "conv2d_mul", "bias_add", "batch_norm",
           "max_pool2d", "leaky_relu", "dnnb_net_create", "dnnb_net_add_block", "dnnb_net_finalize",
           "dnnb_net_set_latency", "dnnb_net_set_latency_widths", "dnnb_net_knobs",
           "dnnb_net_num_blocks", "dnnb_net_out_shape", "dnnb_net_forward", "dnnb_net_forward_async",
           "dnnb_net_forward_device", "dnnb_net_forward_u8_async", "dnnb_net_forward_u8_device", "dnnb_net_sync",
           "dnnb_net_next_slot", "dnnb_net_wait_slot", "dnnb_net_block_output", "dnnb_net_block_info", "dnnb_net_debug_prepool",
           "dnnb_net_time_blocks", "dnnb_net_launches_per_forward", "dnnb_net_debug_block_limit", "dnnb_net_debug_timeline", "dnnb_net_time_forward",
           "dnnb_net_profile_block", "dnnb_net_profile_read", "dnnb_net_destroy", "dnnb_host_alloc",
           "dnnb_host_free", "dnnb_measure_mma_peak", "dnnb_measure_mma_peak_sustained",
           "dnnb_postprocess", "dnnb_postprocess_candidates", "dnnb_nms", "dnnb_net_set_detect",
           "dnnb_net_detect_async", "dnnb_net_detect_u8_async", "dnnb_net_detect_device",
           "dnnb_net_forward_frames_async", "dnnb_net_detect_frames_async", "dnnb_resize_input"]

# struct dnnb_det { int cls, left, top, right, bottom; float score; }
DET_DTYPE = np.dtype([("cls", np.int32), ("left", np.int32), ("top", np.int32), ("right", np.int32),
                      ("bottom", np.int32), ("score", np.float32)])
CANDIDATES = 845


def check(status):
    if status != 0:
        raise RuntimeError("libdnn_b200: " + lib.dnnb_last_error().decode("utf-8", "replace"))


def fptr(a):
    return a.ctypes.data_as(FP)


def f32c(a):
    """float32, C-contiguous view/copy (the reference hands numpy buffers to C as float*)."""
    return np.ascontiguousarray(a, dtype=np.float32)


def require_buffer(a, dtype, shape, what):
    """The async entry points hand raw pointers to cudaMemcpyAsync: anything but a C-contiguous array of exactly
    the expected dtype and shape would be read or written out of bounds, so it is rejected here."""
    if not isinstance(a, np.ndarray):
        raise TypeError("%s must be a numpy array, got %s" % (what, type(a).__name__))
    if a.dtype != np.dtype(dtype):
        raise TypeError("%s must be %s, got %s" % (what, np.dtype(dtype), a.dtype))
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("%s must have shape %s, got %s" % (what, tuple(shape), tuple(a.shape)))
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("%s must be C-contiguous" % what)
    return a


def device_count():
    n = c_int(0)
    check(lib.dnnb_device_count(ctypes.byref(n)))
    return n.value


class PinnedArray:
    """A numpy array (float32 by default) over page-locked host memory (async H2D/D2H)."""

    def __init__(self, shape, dtype=np.float32):
        self.shape = tuple(int(s) for s in shape)
        nbytes = int(np.prod(self.shape)) * np.dtype(dtype).itemsize
        self._ptr = c_void_p()
        check(lib.dnnb_host_alloc(nbytes, ctypes.byref(self._ptr)))
        buf = (ctypes.c_ubyte * nbytes).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(self.shape)

    def free(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            lib.dnnb_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Net:
    """Resident fused network (dnnb_net_*).  blocks: list of dicts with keys kernel, padding,
    biases, mean, variance, gamma, epsilon, leaky, pool (None or (k, s, padding))."""

    def __init__(self, in_shape, blocks, precision="tf32", device=0, latency=None, latency_widths=None):
        """latency: True / False selects the small-batch latency plan for this network (split-K, PDL chain,
        CUDA-graph replay; include/dnn_b200.h, dnnb_net_finalize) regardless of DNNB_LATENCY_MODE; None leaves
        the choice to the environment.  latency_widths: per-block N-tile widths under that plan (0 = the planner's
        modelled choice; dnnb_net_set_latency_widths, tools/tune_latency.py)."""
        b, h, w, c = (int(v) for v in in_shape)
        self.in_shape = (b, h, w, c)
        self.precision = precision
        self.device = device
        self.latency = latency
        self._h = c_void_p()
        check(lib.dnnb_net_create(device, PRECISIONS[precision], b, h, w, c, ctypes.byref(self._h)))
        try:
            for blk in blocks:
                k = f32c(blk["kernel"])
                kh, kw, ic, oc = k.shape
                opt = lambda key: None if blk.get(key) is None else f32c(blk[key])
                bias, mean, var, gamma = opt("biases"), opt("mean"), opt("variance"), opt("gamma")
                p = lambda a: None if a is None else fptr(a)
                pool = blk.get("pool")
                pk, ps, pp = (0, 0, 1) if pool is None else (pool[0], pool[1], 1 if pool[2] == 'SAME' else 0)
                check(lib.dnnb_net_add_block(self._h, fptr(k), kh, kw, ic, oc, 1 if blk.get("padding", 'SAME') == 'SAME' else 0,
                                             p(bias), p(mean), p(var), p(gamma), float(blk.get("epsilon", 0.0)),
                                             1 if blk.get("leaky") else 0, pk, ps, pp))
            if latency is not None:
                check(lib.dnnb_net_set_latency(self._h, 1 if latency else 0))
            if latency_widths:
                arr = (c_int * len(latency_widths))(*[int(v) for v in latency_widths])
                check(lib.dnnb_net_set_latency_widths(self._h, arr, len(latency_widths)))
            check(lib.dnnb_net_finalize(self._h))
            o = [c_int(0) for _ in range(4)]
            check(lib.dnnb_net_out_shape(self._h, *[ctypes.byref(v) for v in o]))
            self.out_shape = tuple(v.value for v in o)
            n = c_int(0)
            check(lib.dnnb_net_num_blocks(self._h, ctypes.byref(n)))
            self.num_blocks = n.value
        except Exception:
            self.close()
            raise

    def close(self):
        if self._h is not None and self._h.value:
            lib.dnnb_net_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, x, out=None):
        x = f32c(x)
        assert tuple(x.shape) == self.in_shape, (x.shape, self.in_shape)
        if out is None:
            out = np.empty(self.out_shape, dtype=np.float32)
        require_buffer(out, np.float32, self.out_shape, "out")
        check(lib.dnnb_net_forward(self._h, x.ctypes.data, out.ctypes.data))
        return out

    def forward_async(self, x, out):
        """x, out: C-contiguous float32 arrays (ideally PinnedArray.array) kept alive until sync()."""
        require_buffer(x, np.float32, self.in_shape, "x")
        require_buffer(out, np.float32, self.out_shape, "out")
        check(lib.dnnb_net_forward_async(self._h, x.ctypes.data, out.ctypes.data))

    def forward_u8(self, frames, bgr=True, out=None):
        """frames: uint8 [B,H,W,3] (BGR as cv2 delivers them when bgr=True).  x = float32(u8/255.) with the
        channel reversal of proj3/__init__.py:8-12 happens on the device."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        assert tuple(frames.shape) == self.in_shape, (frames.shape, self.in_shape)
        if out is None:
            out = np.empty(self.out_shape, dtype=np.float32)
        self.forward_u8_async(frames, out, bgr)
        self.sync()
        return out

    def forward_frames(self, frames, bgr=True, out=None):
        """frames: uint8 [B,h,w,3] of ANY h, w: cv2.resize to the network size, / 255., BGR -> RGB on
        the device (all of proj3/__init__.py:8-12), then the network."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        assert frames.ndim == 4 and frames.shape[0] == self.in_shape[0] and frames.shape[3] == 3, frames.shape
        if out is None:
            out = np.empty(self.out_shape, dtype=np.float32)
        require_buffer(out, np.float32, self.out_shape, "out")
        check(lib.dnnb_net_forward_frames_async(self._h, frames.ctypes.data, frames.shape[1], frames.shape[2],
                                                1 if bgr else 0, out.ctypes.data))
        self.sync()
        return out

    def forward_u8_async(self, frames, out, bgr=True):
        require_buffer(frames, np.uint8, self.in_shape, "frames")
        require_buffer(out, np.float32, self.out_shape, "out")
        check(lib.dnnb_net_forward_u8_async(self._h, frames.ctypes.data, 1 if bgr else 0, out.ctypes.data))

    def set_detect(self, score_threshold=0.3, iou_threshold=0.3, max_det=32):
        check(lib.dnnb_net_set_detect(self._h, score_threshold, iou_threshold, max_det))
        self.max_det = max_det

    def detect_async(self, x, dets, counts, bgr=True):
        """Forward + on-device box decoding.  x: float32 [B,H,W,C] or uint8 frames; dets: DET_DTYPE
        [B, max_det]; counts: int32 [B] (ideally PinnedArray.array), kept alive until sync()."""
        require_buffer(dets, DET_DTYPE, (self.in_shape[0], getattr(self, "max_det", 32)), "dets")
        require_buffer(counts, np.int32, (self.in_shape[0],), "counts")
        if isinstance(x, np.ndarray) and x.dtype == np.uint8:
            require_buffer(x, np.uint8, None, "frames")
            if x.ndim != 4 or x.shape[0] != self.in_shape[0] or x.shape[3] != 3:
                raise ValueError("frames must be [%d, h, w, 3], got %s" % (self.in_shape[0], tuple(x.shape)))
            check(lib.dnnb_net_detect_frames_async(self._h, x.ctypes.data, x.shape[1], x.shape[2], 1 if bgr else 0,
                                                   dets.ctypes.data, counts.ctypes.data))
        else:
            require_buffer(x, np.float32, self.in_shape, "x")
            check(lib.dnnb_net_detect_async(self._h, x.ctypes.data, dets.ctypes.data, counts.ctypes.data))

    def detect(self, x, bgr=True):
        """-> (dets [B, max_det] DET_DTYPE, counts [B]); only the first min(count, max_det) entries of a row are valid."""
        x = np.ascontiguousarray(x)
        if x.dtype != np.uint8:
            x = f32c(x)
            assert tuple(x.shape) == self.in_shape, (x.shape, self.in_shape)
        else:       # uint8 frames of any size are resized on the device
            assert x.ndim == 4 and x.shape[0] == self.in_shape[0] and x.shape[3] == 3, x.shape
        dets = np.zeros((self.in_shape[0], getattr(self, "max_det", 32)), dtype=DET_DTYPE)
        counts = np.zeros(self.in_shape[0], dtype=np.int32)
        self.detect_async(x, dets, counts, bgr)
        self.sync()
        return dets, counts

    def detect_device(self, in_ptr, dets_ptr, counts_ptr, is_u8=False, bgr=True, stream=None):
        check(lib.dnnb_net_detect_device(self._h, in_ptr, 1 if is_u8 else 0, 1 if bgr else 0, dets_ptr, counts_ptr, stream))

    def forward_device(self, in_ptr, out_ptr=None, stream=None):
        check(lib.dnnb_net_forward_device(self._h, in_ptr, out_ptr, stream))

    def sync(self):
        check(lib.dnnb_net_sync(self._h))

    def next_slot(self):
        """Pipeline slot (0 / 1) the next *_async call will use."""
        n = c_int(0)
        check(lib.dnnb_net_next_slot(self._h, ctypes.byref(n)))
        return n.value

    def wait_slot(self, slot):
        """Blocks until the most recent job queued on `slot` has delivered its results to the host buffers."""
        check(lib.dnnb_net_wait_slot(self._h, slot))

    def block_info(self, i):
        info = (c_int * 16)()
        check(lib.dnnb_net_block_info(self._h, i, info))
        v = list(info)
        return {"kernel": KERNEL_NAMES[v[0]], "out_shape": (self.in_shape[0], v[1], v[2], v[3]),
                "pool_fused": bool(v[4]), "bn_tile": v[5], "swizzle": v[6], "stages": v[7], "grid": v[8],
                "pool_separate": bool(v[9]), "act_bytes": v[10], "launches": v[11], "tiles": v[12],
                "splits": v[13], "cluster_splitk": bool(v[14])}

    def block_output(self, i):
        shape = self.block_info(i)["out_shape"]
        out = np.empty(shape, dtype=np.float32)
        check(lib.dnnb_net_block_output(self._h, i, fptr(out)))
        return out

    def debug_prepool(self, i, conv_hw):
        """Pre-pool activations of a block whose pool is fused (see dnnb_net_debug_prepool)."""
        info = self.block_info(i)
        out = np.empty((self.in_shape[0], conv_hw[0], conv_hw[1], info["out_shape"][3]), dtype=np.float32)
        check(lib.dnnb_net_debug_prepool(self._h, i, fptr(out)))
        return out

    def time_blocks(self, warmup=3, iters=10):
        ms = (c_float * self.num_blocks)()
        check(lib.dnnb_net_time_blocks(self._h, warmup, iters, ms))
        return list(ms)

    def profile_block(self, block):
        check(lib.dnnb_net_profile_block(self._h, block))

    def profile_read(self):
        ms, n = c_float(0), c_int(0)
        check(lib.dnnb_net_profile_read(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def debug_timeline(self, block, read=False):
        """Arms (read=False) or reads (read=True -> uint64 [ctas, TIMELINE_SLOTS] of %globaltimer ns) the per-CTA phase
        stamps of a block (slots: include/dnn_b200.h, dnnb_net_debug_timeline)."""
        n = c_int(0)
        if not read:
            check(lib.dnnb_net_debug_timeline(self._h, block, None, ctypes.byref(n)))
            return n.value
        grid = self.block_info(block)["grid"]
        out = np.zeros((grid, TIMELINE_SLOTS), dtype=np.uint64)
        check(lib.dnnb_net_debug_timeline(self._h, block, out.ctypes.data, ctypes.byref(n)))
        return out[:n.value]

    def time_forward(self, warmup=10, iters=60):
        """Median device time (us) of a forward from the network's own buffers (graph replays under the latency plan)."""
        us = c_float(0)
        check(lib.dnnb_net_time_forward(self._h, warmup, iters, ctypes.byref(us)))
        return us.value

    def debug_block_limit(self, blocks):
        check(lib.dnnb_net_debug_block_limit(self._h, blocks))

    def knobs(self):
        """The DNNB_* switches that were not at their default when this network was planned, as a dict
        (empty for the default plan); an explicit latency= argument shows up as 'set_latency'."""
        buf = ctypes.create_string_buffer(4096)
        check(lib.dnnb_net_knobs(self._h, buf, len(buf)))
        text = buf.value.decode()
        return dict(item.split("=", 1) for item in text.split(",")) if text else {}

    def launches_per_forward(self):
        n = c_int(0)
        check(lib.dnnb_net_launches_per_forward(self._h, ctypes.byref(n)))
        return n.value


def resize_input(frames, hw=416, bgr=True, device=0):
    """proj3/__init__.py:8-12 on the GPU: uint8 [B,h,w,3] (or one [h,w,3]) -> float32 [B,hw,hw,3]."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    single = frames.ndim == 3
    if single:
        frames = frames[None]
    out = np.empty((frames.shape[0], hw, hw, 3), dtype=np.float32)
    check(lib.dnnb_resize_input(device, frames.ctypes.data, frames.shape[0], frames.shape[1], frames.shape[2], hw, hw,
                                1 if bgr else 0, out.ctypes.data))
    return out[0] if single else out


def postprocess(pred, score_threshold=0.3, iou_threshold=0.3, max_det=32, device=0):
    """Head tensors [B,13,13,125] (or one [13,13,125]) -> (dets [B, max_det] DET_DTYPE, counts [B]) on the GPU."""
    pred = f32c(pred).reshape(-1, 13, 13, 125)
    dets = np.zeros((pred.shape[0], max_det), dtype=DET_DTYPE)
    counts = np.zeros(pred.shape[0], dtype=np.int32)
    check(lib.dnnb_postprocess(device, pred.ctypes.data, pred.shape[0], score_threshold, iou_threshold, max_det,
                               dets.ctypes.data, counts.ctypes.data))
    return dets, counts


def postprocess_candidates(pred, device=0):
    """All 845 decoded candidates per image, (row, col, box) order, before the threshold."""
    pred = f32c(pred).reshape(-1, 13, 13, 125)
    cand = np.zeros((pred.shape[0], CANDIDATES), dtype=DET_DTYPE)
    check(lib.dnnb_postprocess_candidates(device, pred.ctypes.data, pred.shape[0], cand.ctypes.data))
    return cand


def nms(cand, score_threshold=0.3, iou_threshold=0.3, max_det=32, device=0):
    """Threshold + stable sort + greedy NMS alone, on candidates [B,845] DET_DTYPE."""
    cand = np.ascontiguousarray(cand, dtype=DET_DTYPE).reshape(-1, CANDIDATES)
    dets = np.zeros((cand.shape[0], max_det), dtype=DET_DTYPE)
    counts = np.zeros(cand.shape[0], dtype=np.int32)
    check(lib.dnnb_nms(device, cand.ctypes.data, cand.shape[0], score_threshold, iou_threshold, max_det,
                       dets.ctypes.data, counts.ctypes.data))
    return dets, counts


def measure_mma_peak(precision="tf32", device=0, seconds=0.0):
    """tcgen05 issue-loop throughput in TFLOP/s: burst (seconds == 0) or sustained over `seconds`."""
    t = c_float(0)
    if seconds > 0:
        check(lib.dnnb_measure_mma_peak_sustained(device, PRECISIONS[precision], float(seconds), ctypes.byref(t)))
    else:
        check(lib.dnnb_measure_mma_peak(device, PRECISIONS[precision], ctypes.byref(t)))
    return t.value


def tune_latency_widths(in_shape, blocks, precision="tf32", device=0, iters=60, sweeps=2, margin_us=0.5, log=None):
    """Measures the N-tile width of every generic-kernel block of a latency-plan network on this device: coordinate
    descent over the blocks, every candidate a freshly planned network timed over whole forwards (PDL chain + graph
    replay), because a block's best width depends on its neighbours (a grid that fills every SM keeps its successor's
    prologue from starting early) and on the part.  Returns {"widths", "tuned_us", "modelled_us"}; pass widths as
    Net(latency_widths=...) / dnnb_net_set_latency_widths to pin the plan (results of different plans differ by fp32
    summation order where the number of K slices differs)."""
    def measure(widths):
        net = Net(in_shape, blocks, precision=precision, device=device, latency=True, latency_widths=widths)
        try:
            info = [net.block_info(i) for i in range(net.num_blocks)]
            return net.time_forward(10, iters), info
        finally:
            net.close()

    n = len(blocks)
    widths = [0] * n
    base, info0 = measure(widths)
    best = base
    tunable = [i for i in range(n) if info0[i]["kernel"] in ("umma_im2col", "umma_tiled_pool")]
    for sweep in range(sweeps):
        changed = False
        for i in tunable:
            for cand in (256, 128, 64, 32):
                if cand == (widths[i] or info0[i]["bn_tile"]):
                    continue
                trial = list(widths)
                trial[i] = cand
                t, info = measure(trial)
                if info[i]["bn_tile"] != cand or info[i]["kernel"] == "simt":
                    continue                            # no kernel of that width: the planner fell back
                if t < best - margin_us:
                    t2, _ = measure(trial)              # a second look before a plan is kept
                    if max(t, t2) < best - margin_us:
                        best, widths, changed = max(t, t2), trial, True
                if log:
                    log("sweep %d block %d width %d: %.1f us (best %.1f)" % (sweep, i, cand, t, best))
        if not changed:
            break
    return {"widths": widths, "tuned_us": best, "modelled_us": base}
