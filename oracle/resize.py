"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's input preprocessing,
`resize_input` (cs492-projects/proj3/__init__.py:8-12):

    imsz = cv2.resize(im, (416, 416))      # uint8 BGR frame, default INTER_LINEAR
    imsz = imsz / 255.                      # float64
    imsz = imsz[:, :, ::-1]                 # BGR -> RGB
    return np.asarray(imsz, dtype=np.float32)

The bilinear resize lives in a third-party dependency that is not under /root/reference:
OpenCV (the reference pins no version; this image has opencv-python 4.13.0).  Its published
8-bit INTER_LINEAR algorithm (modules/imgproc/src/resize.cpp: the coefficient set-up of
cv::resize, HResizeLinear, VResizeLinear<uchar, int, short, FixedPtCast<...>>) is restated below:
  scale = 1 / (dst / src)                                   (double)
  f = float((d + 0.5) * scale - 0.5); s = floor(f); f -= s  (float32)
  x only: s < 0 -> (s, f) = (0, 0);  s >= src-1 -> (s, f) = (src-1, 0)
  y: the two source rows s, s+1 are clipped to [0, src-1]; the weights are NOT reset
  weights: short(round_half_even((1 - f) * 2048)), short(round_half_even(f * 2048))
  horizontal: H = S[s] * a0 + S[s+1] * a1                   (int32, scale 2^11)
  vertical:   out = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2
Pinned: oracle/make_golden.py stores cv2.resize outputs in tests/golden/resize.npz and
tests/test_resize_oracle.py checks this restatement against them bit for bit (and against cv2
itself when it is importable).  Only tests/, smoke() and bench.py's baseline legs import this.
"""
import math

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def linear_coeffs(src, dst, clamp):
    """(offset int64 [dst], weights int64 [dst,2]) of one axis; `clamp` = the x-axis border rule."""
    scale = 1.0 / (float(dst) / float(src))
    ofs = np.zeros(dst, dtype=np.int64)
    w = np.zeros((dst, 2), dtype=np.int64)
    f32 = np.float32
    for d in range(dst):
        f = f32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = f32(f - f32(s))
        if clamp:
            if s < 0:
                s, f = 0, f32(0)
            if s >= src - 1:
                s, f = src - 1, f32(0)
        ofs[d] = s
        w[d, 0] = int(np.rint(f32((f32(1) - f) * f32(COEF_SCALE))))
        w[d, 1] = int(np.rint(f32(f * f32(COEF_SCALE))))
    return ofs, w


def resize_bilinear_u8(img, dst_h, dst_w):
    """uint8 [H,W,C] -> uint8 [dst_h,dst_w,C], cv2.resize(img, (dst_w, dst_h)) bit for bit."""
    img = np.asarray(img, dtype=np.uint8)
    sh, sw = img.shape[:2]
    xo, xa = linear_coeffs(sw, dst_w, True)
    yo, ya = linear_coeffs(sh, dst_h, False)
    src = img.astype(np.int64)
    x1 = np.minimum(xo + 1, sw - 1)
    hor = src[:, xo] * xa[:, 0][None, :, None] + src[:, x1] * xa[:, 1][None, :, None]
    r0, r1 = hor[np.clip(yo, 0, sh - 1)], hor[np.clip(yo + 1, 0, sh - 1)]
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def resize_input(frame, hw=416):
    """The whole of proj3/__init__.py:8-12 on one uint8 BGR frame -> float32 RGB [hw,hw,3]."""
    imsz = resize_bilinear_u8(frame, hw, hw)
    imsz = imsz / 255.
    imsz = imsz[:, :, ::-1]
    return np.asarray(imsz, dtype=np.float32)
