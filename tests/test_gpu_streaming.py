"""GPU test of the streaming loop (proj1/__init__.py:71-137): frames in -> label boxes out,
pipelined, against oracle/resize.py (cv2-exact resize_input) + oracle/postproc.py (the reference's
postprocessing) around this engine's own forward."""
import numpy as np
import pytest

import _lib
import postproc as pp
import resize as rs
from streaming import StreamingDetector, yolo_spec
from synth import make_weights, tame_head

pytestmark = pytest.mark.gpu


def _as_rows(boxes):
    from detect import CLASS_NAMES
    return [(CLASS_NAMES.index(n), lt[0], lt[1], rb[0], rb[1]) for n, lt, rb, _ in boxes]


@pytest.mark.parametrize("batch,n_frames", [(1, 5), (4, 10)])
def test_streaming_detector_equals_frame_by_frame_reference_pipeline(batch, n_frames):
    weights = tame_head(make_weights(0))
    h, w = 360, 480
    rng = np.random.default_rng(21)
    frames = rng.integers(0, 256, (n_frames, h, w, 3), dtype=np.uint8)
    det = StreamingDetector(weights, (h, w), batch=batch, max_det=64)
    got = list(det.run(iter(frames)))                  # ragged last job at batch 4: 10 = 4 + 4 + 2
    assert len(got) == n_frames and det.frames_done == n_frames and det.fps() > 0
    # reference pipeline, frame by frame: oracle resize -> this engine's forward (same plan) -> oracle postprocessing
    ref_net = _lib.Net((batch, 416, 416, 3), yolo_spec(weights), precision="tf32", latency=True if batch == 1 else None)
    checked = 0
    for j0 in range(0, n_frames, batch):
        chunk = frames[j0:j0 + batch]
        x = np.zeros((batch, 416, 416, 3), np.float32)
        x[:len(chunk)] = np.stack([rs.resize_input(f, 416) for f in chunk])
        if len(chunk) < batch:      # the detector pads a short job with black frames
            x[len(chunk):] = rs.resize_input(np.zeros((h, w, 3), np.uint8), 416)
        head = ref_net.forward(x)
        stable = pp.stable_images(head)
        for i in range(len(chunk)):
            if stable[i]:
                want = pp.detect_image(head[i], zero_union="drop")
                assert _as_rows(got[j0 + i]) == [t[:5] for t in want[:64]], "frame %d" % (j0 + i)
                checked += 1
    assert checked >= 2           # random frames: only images whose every decision has a float64 margin are compared exactly
    ref_net.close()
    # rescale=True maps boxes to the frame like proj1's re_resize_pos
    det2 = StreamingDetector(weights, (h, w), batch=batch, max_det=64, rescale=True)
    got2 = list(det2.run(iter(frames[:batch])))
    for a, b in zip(got[:batch], got2):
        assert [(n, (int(lt[0] * w / 416), int(lt[1] * h / 416)), (int(rb[0] * w / 416), int(rb[1] * h / 416)), c)
                for n, lt, rb, c in a] == b
    det.close(); det2.close()


def test_streaming_detector_rejects_wrong_frames():
    det = StreamingDetector(tame_head(make_weights(0)), (64, 64), batch=2)
    with pytest.raises(ValueError):
        det.submit(np.zeros((2, 64, 64, 3), np.float32))
    with pytest.raises(ValueError):
        det.submit(np.zeros((3, 64, 64, 3), np.uint8))
    with pytest.raises(ValueError):
        det.submit(np.zeros((2, 32, 64, 3), np.uint8))
    det.close()
