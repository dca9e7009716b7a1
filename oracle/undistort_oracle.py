"""CPU oracle of the frame ingest with lens undistortion.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference undistorts every camera frame with
``cv.undistort(frame, mtx, dist, None, new_camera_matrix)`` and crops it to the ROI of
``cv.getOptimalNewCameraMatrix`` (detect_pose.py:147-183, called from process_frame, detect_pose.py:611-619)
before converting it to gray (detect_pose.py:602).  cv2 is a binary here, so this file restates what
cv::undistort does for 8-bit images - published OpenCV 4.x algorithm, checked bit for bit against the installed
cv2 4.13 by tests/test_cpu_oracle.py:

* maps: stripes of ``max(1, 4096 // width)`` rows, each with the principal point of the new camera matrix shifted
  by the stripe origin; normalised coordinates through the inverse new camera matrix, Brown-Conrady distortion
  (k1 k2 p1 p2 k3) in float64, source coordinates in 1/32 px fixed point (round half to even);
* remap: INTER_LINEAR with the 32 x 32 table of int16 weights (float32 products scaled by 2^15, saturated, the
  rounding remainder folded into one weight), ``(sum + 2^14) >> 15``, BORDER_CONSTANT 0.

The quantisation to 1/32 px makes the result insensitive to the last bits of the float64 arithmetic (a different
operation order changes a pixel only if a coordinate lies within ~1e-12 of a rounding boundary).
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
COEF_SCALE = 1 << 15


def bilinear_table() -> np.ndarray:
    """[1024, 4] int32 weights (w00, w01, w10, w11) indexed by (fy * 32 + fx), as cv::initInterTab2D builds them."""
    t1 = np.zeros((INTER_TAB_SIZE, 2), np.float32)
    for i in range(INTER_TAB_SIZE):
        x = np.float32(i) * np.float32(1.0 / INTER_TAB_SIZE)
        t1[i] = (np.float32(1.0) - x, x)
    flat = np.zeros(INTER_TAB_SIZE * INTER_TAB_SIZE * 4 + 8, np.int32)      # room for the look-ahead of the fix-up loop
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            base = (i * INTER_TAB_SIZE + j) * 4
            isum = 0
            for k1 in range(2):
                for k2 in range(2):
                    v = np.float32(t1[i, k1] * t1[j, k2])
                    q = int(np.rint(np.float64(v) * COEF_SCALE))
                    q = max(-32768, min(32767, q))
                    flat[base + k1 * 2 + k2] = q
                    isum += q
            if isum != COEF_SCALE:
                diff = isum - COEF_SCALE
                big = small = base + 3
                for k1 in (1, 2):
                    for k2 in (1, 2):
                        idx = base + k1 * 2 + k2
                        if flat[idx] < flat[small]:
                            small = idx
                        elif flat[idx] > flat[big]:
                            big = idx
                if diff < 0:
                    flat[big] -= diff
                else:
                    flat[small] -= diff
    return flat[:INTER_TAB_SIZE * INTER_TAB_SIZE * 4].reshape(-1, 4)


def fixed_point_maps(mtx, dist, new_mtx, width: int, height: int):
    """Source coordinates in 1/32 px of every pixel of the undistorted (uncropped) frame: (iu, iv) int64 [H, W]."""
    k = np.zeros(5)
    d = np.asarray(dist, np.float64).ravel()
    k[:min(len(d), 5)] = d[:5]
    k1, k2, p1, p2, k3 = k
    mtx = np.asarray(mtx, np.float64)
    new_mtx = np.asarray(new_mtx, np.float64)
    fx, fy, u0, v0 = mtx[0, 0], mtx[1, 1], mtx[0, 2], mtx[1, 2]
    stripe = min(max(1, (1 << 12) // max(width, 1)), height)
    iu = np.zeros((height, width), np.int64)
    iv = np.zeros((height, width), np.int64)
    j = np.arange(width, dtype=np.float64)
    for y0 in range(0, height, stripe):
        ar = new_mtx.copy()
        ar[1, 2] = new_mtx[1, 2] - y0
        ir = np.linalg.inv(ar).ravel()
        for i in range(min(stripe, height - y0)):
            _x = i * ir[1] + ir[2] + j * ir[0]
            _y = i * ir[4] + ir[5] + j * ir[3]
            _w = i * ir[7] + ir[8] + j * ir[6]
            w = 1.0 / _w
            x, y = _x * w, _y * w
            x2, y2 = x * x, y * y
            r2, _2xy = x2 + y2, 2 * x * y
            kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
            xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2)
            yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy
            iu[y0 + i] = np.rint((fx * xd + u0) * INTER_TAB_SIZE).astype(np.int64)
            iv[y0 + i] = np.rint((fy * yd + v0) * INTER_TAB_SIZE).astype(np.int64)
    return iu, iv


def remap_linear_u8(src: np.ndarray, iu: np.ndarray, iv: np.ndarray, table: np.ndarray) -> np.ndarray:
    """cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) of an 8-bit image with CV_16SC2 fixed-point maps."""
    sh, sw = src.shape[:2]
    wrap = lambda a: (a + 32768) % 65536 - 32768                # the (short) cast of the integer part
    sx, sy = wrap(iu >> INTER_BITS), wrap(iv >> INTER_BITS)
    wts = table[(iv & (INTER_TAB_SIZE - 1)) * INTER_TAB_SIZE + (iu & (INTER_TAB_SIZE - 1))].astype(np.int64)

    def px(yy, xx):
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        v = src[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)].astype(np.int64)
        return np.where(ok[..., None] if src.ndim == 3 else ok, v, 0)

    acc = 0
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        wk = wts[..., k]
        acc = acc + (wk[..., None] if src.ndim == 3 else wk) * px(sy + dy, sx + dx)
    return ((acc + (1 << 14)) >> 15).clip(0, 255).astype(np.uint8)


def undistort(src: np.ndarray, mtx, dist, new_mtx) -> np.ndarray:
    """cv.undistort(src, mtx, dist, None, new_mtx) for 8-bit images of 1 or 3 channels."""
    h, w = src.shape[:2]
    iu, iv = fixed_point_maps(mtx, dist, new_mtx, w, h)
    return remap_linear_u8(src, iu, iv, bilinear_table())


def bgr_to_gray(bgr: np.ndarray) -> np.ndarray:
    """cv.cvtColor(bgr, COLOR_BGR2GRAY) for 8-bit images (fixed point, 15 bits)."""
    b, g, r = (bgr[..., c].astype(np.int64) for c in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


def undistort_frame_gray(frame: np.ndarray, mtx, dist, new_mtx, roi) -> np.ndarray:
    """The reference's ingest: undistort_frame (detect_pose.py:147-183) followed by BGR2GRAY (detect_pose.py:602)."""
    x, y, w, h = (int(v) for v in roi)
    out = undistort(frame, mtx, dist, new_mtx)[y:y + h, x:x + w]
    return bgr_to_gray(out) if out.ndim == 3 else out
