"""Frames per second of the streaming loop (streaming.StreamingDetector; proj1/__init__.py:71-137 with the device doing
resize_input + the network + postprocessing): synthetic BGR frames of a video-like size in, label boxes out, wall clock
around the whole loop including the host-side frame copies.   python tools/stream_fps.py [h w] > result.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dnn-inference-engine_b200"), os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402

from streaming import StreamingDetector  # noqa: E402
from synth import make_weights, tame_head  # noqa: E402

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (720, 1280)
weights = tame_head(make_weights(0))
rng = np.random.default_rng(3)
pool = rng.integers(0, 256, (64, h, w, 3), dtype=np.uint8)          # 64 distinct frames, cycled
out = {"frame_hw": [h, w], "data": "synthetic uint8 BGR frames", "runs": []}
for batch, n_frames in ((1, 2000), (8, 8000), (64, 32000)):
    for precision in ("tf32", "bf16"):
        det = StreamingDetector(weights, (h, w), batch=batch, precision=precision, max_det=32)
        for _ in det.run(pool[i % 64] for i in range(4 * batch)):       # warm-up
            pass
        det.frames_done, det.seconds = 0, 0.0
        t0 = time.perf_counter()
        boxes = 0
        for b in det.run(pool[i % 64] for i in range(n_frames)):
            boxes += len(b)
        dt = time.perf_counter() - t0
        out["runs"].append({"batch": batch, "precision": precision, "frames": n_frames, "fps": round(n_frames / dt, 1),
                            "ms_per_frame": round(1e3 * dt / n_frames, 4), "boxes_per_frame": round(boxes / n_frames, 2)})
        print(out["runs"][-1], file=sys.stderr, flush=True)
        det.close()
print(json.dumps(out, indent=1))
