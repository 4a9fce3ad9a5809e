"""streaming -- frames in, label boxes out, continuously: the loop of proj1/__init__.py:71-137
(`video_object_detection`: read frame -> resize_input -> inference -> postprocessing -> draw) with everything
between "frame read" and "boxes" on the GPU and pipelined.

    det = StreamingDetector(weights, frame_hw=(480, 640), batch=8)
    for boxes in det.run(frame_iterator):       # one list of label boxes per frame, in frame order
        ...                                     # [(class name, (left, top), (right, bottom), colour)] like the reference

One job = `batch` uint8 BGR frames: H2D of the raw frames (3 bytes per pixel), cv2-exact bilinear resize to the
network size, /255, BGR->RGB, the 9 fused conv blocks, box decoding + NMS, D2H of the boxes only
(dnnb_net_detect_frames_async).  The library double-buffers its device-side staging, this class double-buffers
the page-locked host side, so the copies of job k+1 and the results of job k-1 overlap the kernels of job k; the
loop only ever waits for the job before the one it has just queued (dnnb_net_wait_slot).
Box coordinates are those of the 416x416 network input, as `postprocessing` returns them; `rescale=True` maps
them to the frame the way proj1's draw_output_frame does (re_resize_pos, proj1/__init__.py:52-56).
There is no CPU path."""
import time

import numpy as np

import _lib
from detect import label_boxes
from synth import POOLS


def yolo_spec(weights, epsilon=1e-5):
    """Block list for _lib.Net from the reference's weight container (proj3/yolov2tiny.py:28-79)."""
    spec = []
    n = len(weights)
    for i, w in enumerate(weights):
        blk = {"kernel": w["kernel"], "padding": "SAME", "biases": w["biases"]}
        if "gamma" in w:
            blk.update(mean=w["moving_mean"], variance=w["moving_variance"], gamma=w["gamma"], epsilon=epsilon, leaky=True)
        pool = POOLS[i] if n == len(POOLS) else None
        blk["pool"] = None if pool is None else (pool[0], pool[1], "SAME")
        spec.append(blk)
    return spec


class StreamingDetector:
    def __init__(self, weights, frame_hw, batch=1, precision="tf32", device=0, max_det=32, bgr=True, in_hw=416,
                 score_threshold=0.3, iou_threshold=0.3, latency=None, rescale=False):
        self.batch, self.frame_hw, self.bgr, self.max_det, self.in_hw = int(batch), tuple(frame_hw), bgr, max_det, in_hw
        self.rescale = rescale
        self.net = _lib.Net((self.batch, in_hw, in_hw, 3), yolo_spec(weights), precision=precision, device=device,
                            latency=latency if latency is not None else (True if self.batch == 1 else None))
        self.net.set_detect(score_threshold, iou_threshold, max_det)
        h, w = self.frame_hw
        self._frames = [_lib.PinnedArray((self.batch, h, w, 3), dtype=np.uint8) for _ in range(2)]
        self._dets = [_lib.PinnedArray((self.batch, max_det), dtype=_lib.DET_DTYPE) for _ in range(2)]
        self._counts = [_lib.PinnedArray((self.batch,), dtype=np.int32) for _ in range(2)]
        self._pending = []           # (slot, number of real frames) of the jobs in flight, oldest first
        self.frames_done = 0
        self.seconds = 0.0

    # -- one job ----------------------------------------------------------------------------------
    def _queue(self, slot, n):
        buf = self._frames[slot].array
        if n < self.batch:
            buf[n:] = 0                                   # a short last job is padded with black frames
        self.net.detect_async(buf, self._dets[slot].array, self._counts[slot].array, self.bgr)
        self._pending.append((slot, n))

    def submit(self, frames):
        """Queues one job of up to `batch` frames ([n, h, w, 3] uint8; a short last job is padded).  At most two
        jobs are in flight: when both slots are busy this first collects (and returns) the older job's boxes."""
        frames = np.asarray(frames)
        if frames.dtype != np.uint8 or frames.ndim != 4 or tuple(frames.shape[1:]) != self.frame_hw + (3,) or not (1 <= frames.shape[0] <= self.batch):
            raise ValueError("frames must be uint8 [1..%d, %d, %d, 3], got %s %s" % (self.batch, self.frame_hw[0], self.frame_hw[1], frames.dtype, frames.shape))
        done = self.collect() if len(self._pending) == 2 else None
        slot = self.net.next_slot()
        self._frames[slot].array[:frames.shape[0]] = frames
        self._queue(slot, frames.shape[0])
        return done

    def collect(self):
        """Waits for the oldest job in flight and returns its label boxes, one list per frame."""
        if not self._pending:
            return []
        slot, n = self._pending.pop(0)
        self.net.wait_slot(slot)
        dets, counts = self._dets[slot].array, self._counts[slot].array
        out = [label_boxes(dets[i], counts[i]) for i in range(n)]
        if self.rescale:
            h, w = self.frame_hw
            re = lambda pos: (int(pos[0] * w / self.in_hw), int(pos[1] * h / self.in_hw))
            out = [[(name, re(lt), re(rb), col) for name, lt, rb, col in boxes] for boxes in out]
        self.frames_done += n
        return out

    # -- the loop ---------------------------------------------------------------------------------
    def run(self, frames):
        """frames: iterable of [h, w, 3] uint8 frames (cv2.VideoCapture.read order).  Yields every frame's label
        boxes, in order, while later frames are already in flight.  Each frame is copied once, straight into the
        page-locked buffer of the pipeline slot its job will use."""
        t0 = time.perf_counter()
        it = iter(frames)
        shape = self.frame_hw + (3,)
        while True:
            if len(self._pending) == 2:                   # the slot about to be filled still belongs to the oldest job
                for boxes in self.collect():
                    yield boxes
            slot = self.net.next_slot()
            buf = self._frames[slot].array
            n = 0
            for f in it:
                if f.dtype != np.uint8 or f.shape != shape:
                    raise ValueError("frames must be uint8 %s, got %s %s" % (shape, f.dtype, f.shape))
                buf[n] = f
                n += 1
                if n == self.batch:
                    break
            if n == 0:
                break
            self._queue(slot, n)
        while self._pending:
            for boxes in self.collect():
                yield boxes
        self.seconds += time.perf_counter() - t0

    def fps(self):
        return self.frames_done / self.seconds if self.seconds > 0 else 0.0

    def close(self):
        if self.net is not None:
            self.net.sync()
            self.net.close()
            self.net = None
        for group in (self._frames, self._dets, self._counts):
            for p in group:
                p.free()
