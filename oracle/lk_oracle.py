"""CPU oracle for the image pyramid and pyramidal Lucas-Kanade stages.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The mounted reference snapshot
contains no pyramid / optical-flow code (finding 1); the
frozen specification is the behaviour of OpenCV 4.13 (opencv-python-headless
4.13.0.92, third-party binary in this image, source not on disk):

  * ``cv2.pyrDown``  - 5x5 [1 4 6 4 1]^2 / 256 with round-to-nearest, BORDER_REFLECT_101
  * ``cv2.Scharr(.., CV_16S)`` - [3 10 3] x [-1 0 1], unnormalised, BORDER_REFLECT_101
  * ``cv2.calcOpticalFlowPyrLK(prev, next, pts, None)`` with its defaults
    winSize (21,21), maxLevel 3, criteria (COUNT+EPS, 30, 0.01), minEig 1e-4.

``*_cv`` functions call OpenCV (the oracle proper).  ``*_np`` functions are a
numpy restatement of the published algorithm (OpenCV modules/video
lkpyramid.cpp, modules/imgproc pyramids.cpp / deriv.cpp, restated from their
documented behaviour) that tests pin against the ``*_cv`` functions; they
exist so the semantics the CUDA kernels implement are written down in one
readable place.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import cv2 as cv
import numpy as np

WIN = 21
HALF_WIN = 10.0
MAX_LEVEL = 3
MAX_ITERS = 30
EPS_SQ = 0.01 * 0.01
MIN_EIG = 1e-4
W_BITS = 14
FLT_SCALE = np.float32(1.0 / (1 << 20))


# ----------------------------------------------------------------------------
# OpenCV (the oracle proper)
# ----------------------------------------------------------------------------
def pyramid_cv(img: np.ndarray, levels: int = MAX_LEVEL + 1) -> List[np.ndarray]:
    out = [img]
    for _ in range(levels - 1):
        out.append(cv.pyrDown(out[-1]))
    return out


def scharr_cv(img: np.ndarray) -> np.ndarray:
    """(H,W,2) int16, interleaved (dx,dy) as buildOpticalFlowPyramid stores it."""
    return np.stack([cv.Scharr(img, cv.CV_16S, 1, 0), cv.Scharr(img, cv.CV_16S, 0, 1)], axis=-1)


def lk_cv(prev: np.ndarray, nxt: np.ndarray, pts: np.ndarray, threads: int = 0):
    """-> (next_pts (P,2) f32, status (P,) u8, err (P,) f32)."""
    if threads:
        cv.setNumThreads(threads)
    p = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 1, 2)
    nxt_pts, st, err = cv.calcOpticalFlowPyrLK(prev, nxt, p, None)
    return nxt_pts.reshape(-1, 2), st.reshape(-1), err.reshape(-1)


# ----------------------------------------------------------------------------
# numpy restatement
# ----------------------------------------------------------------------------
def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def pyr_down_np(img: np.ndarray) -> np.ndarray:
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    k = np.array([1, 4, 6, 4, 1], dtype=np.int32)
    src = img.astype(np.int32)
    cols = _reflect101(2 * np.arange(ow)[:, None] + np.arange(-2, 3)[None, :], w)      # (ow,5)
    rows = _reflect101(2 * np.arange(oh)[:, None] + np.arange(-2, 3)[None, :], h)      # (oh,5)
    hor = (src[:, cols] * k).sum(axis=-1)                                               # (h,ow)
    ver = (hor[rows, :] * k[None, :, None]).sum(axis=1)                                 # (oh,ow)
    return ((ver + 128) >> 8).astype(np.uint8)


def scharr_np(img: np.ndarray) -> np.ndarray:
    h, w = img.shape
    s = img.astype(np.int32)
    r = _reflect101(np.arange(-1, h + 1), h)
    c = _reflect101(np.arange(-1, w + 1), w)
    p = s[r][:, c]                                                                      # (h+2,w+2)
    sm_v = 3 * p[:-2, :] + 10 * p[1:-1, :] + 3 * p[2:, :]                               # smooth along y
    sm_h = 3 * p[:, :-2] + 10 * p[:, 1:-1] + 3 * p[:, 2:]                               # smooth along x
    dx = sm_v[:, 2:] - sm_v[:, :-2]
    dy = sm_h[2:, :] - sm_h[:-2, :]
    return np.stack([dx, dy], axis=-1).astype(np.int16)


def _descale(v, n):
    return (v + (1 << (n - 1))) >> n


def _cv_round(x: float) -> int:
    # cvRound: round half to even (SSE cvtsd2si)
    return int(np.rint(x))


def _weights(a: np.float32, b: np.float32):
    one = np.float32(1.0)
    s = np.float32(1 << W_BITS)
    iw00 = _cv_round((one - a) * (one - b) * s)
    iw01 = _cv_round(a * (one - b) * s)
    iw10 = _cv_round((one - a) * b * s)
    return iw00, iw01, iw10, (1 << W_BITS) - iw00 - iw01 - iw10


def _pad_img(img):
    return cv.copyMakeBorder(img, WIN, WIN, WIN, WIN, cv.BORDER_REFLECT_101).astype(np.int32)


def _pad_deriv(d):
    return cv.copyMakeBorder(d, WIN, WIN, WIN, WIN, cv.BORDER_CONSTANT, value=0).astype(np.int32)



def _seq_sum_f32(values: np.ndarray) -> np.float32:
    """Left-to-right float32 sum (one rounding per addition), as a scalar `acc += v` loop in C produces."""
    v = np.ascontiguousarray(values, dtype=np.float32).ravel()
    return np.add.accumulate(v, dtype=np.float32)[-1] if v.size else np.float32(0)


SIMD_W = (WIN // 8) * 8          # 16: columns handled eight at a time by OpenCV's 128-bit loop; 16..20 are the scalar tail


def _tensor_sum_f32(p: np.ndarray) -> np.float32:
    """Float32 sum of a (21,21) integer product plane in the order lkpyramid.cpp adds it up (SSE build):
    four float lanes take columns k, k+4, k+8, k+12 of every row (one add each, rows top to bottom), the
    lanes are reduced as (l0+l2)+(l1+l3), columns 16..20 are added one by one into a scalar float, and the
    lane sum is added to that scalar last."""
    f = p.astype(np.float32)                                   # products < 2^24: exact
    lanes = [_seq_sum_f32(f[:, k:SIMD_W:4]) for k in range(4)]
    simd = np.float32(np.float32(lanes[0] + lanes[2]) + np.float32(lanes[1] + lanes[3]))
    return np.float32(_seq_sum_f32(f[:, SIMD_W:]) + simd)


def _mismatch_sum_f32(p: np.ndarray) -> np.float32:
    """Float32 sum of diff*dI over the window in lkpyramid.cpp's order: within each group of eight columns the
    products of columns (k, k+4) are added exactly (pmaddwd, int32), converted to float and added to lane k's
    accumulator (group 0 then group 1, rows top to bottom); result = tail + ((l0+l2)+(l1+l3)), the tail being the
    one-by-one float sum of columns 16..20."""
    pairs = (p[:, 0:SIMD_W].reshape(WIN, SIMD_W // 8, 2, 4).sum(axis=2)).astype(np.float32)   # (21, 2, 4)
    lanes = [_seq_sum_f32(pairs[:, :, k]) for k in range(4)]
    simd = np.float32(np.float32(lanes[0] + lanes[2]) + np.float32(lanes[1] + lanes[3]))
    return np.float32(_seq_sum_f32(p[:, SIMD_W:].astype(np.float32)) + simd)


def lk_np(prev: np.ndarray, nxt: np.ndarray, pts: np.ndarray, levels: int = MAX_LEVEL + 1):
    """Restatement of calcOpticalFlowPyrLK defaults, one point at a time."""
    f32 = np.float32
    pyr_i = [_pad_img(l) for l in pyramid_cv(prev, levels)]
    pyr_d = [_pad_deriv(scharr_cv(l)) for l in pyramid_cv(prev, levels)]
    pyr_j = [_pad_img(l) for l in pyramid_cv(nxt, levels)]
    n = len(pts)
    out = np.zeros((n, 2), f32)
    status = np.ones(n, np.uint8)
    err = np.zeros(n, f32)
    for p in range(n):
        nx = ny = f32(0)
        for level in range(levels - 1, -1, -1):
            rows, cols = pyr_i[level].shape[0] - 2 * WIN, pyr_i[level].shape[1] - 2 * WIN
            sc = f32(1.0 / (1 << level))
            px, py = f32(pts[p, 0]) * sc, f32(pts[p, 1]) * sc
            if level == levels - 1:
                nx, ny = px, py
            else:
                nx, ny = nx * f32(2), ny * f32(2)
            out[p] = (nx, ny)
            px, py = px - f32(HALF_WIN), py - f32(HALF_WIN)
            ix, iy = int(math.floor(px)), int(math.floor(py))
            if ix < -WIN or ix >= cols or iy < -WIN or iy >= rows:
                if level == 0:
                    status[p] = 0
                    err[p] = 0
                continue
            w00, w01, w10, w11 = _weights(px - f32(ix), py - f32(iy))
            y0, x0 = iy + WIN, ix + WIN
            si = pyr_i[level]
            sd = pyr_d[level]
            tmpl = _descale(si[y0:y0 + WIN, x0:x0 + WIN] * w00 + si[y0:y0 + WIN, x0 + 1:x0 + WIN + 1] * w01
                            + si[y0 + 1:y0 + WIN + 1, x0:x0 + WIN] * w10 + si[y0 + 1:y0 + WIN + 1, x0 + 1:x0 + WIN + 1] * w11,
                            W_BITS - 5)
            dt = _descale(sd[y0:y0 + WIN, x0:x0 + WIN] * w00 + sd[y0:y0 + WIN, x0 + 1:x0 + WIN + 1] * w01
                          + sd[y0 + 1:y0 + WIN + 1, x0:x0 + WIN] * w10 + sd[y0 + 1:y0 + WIN + 1, x0 + 1:x0 + WIN + 1] * w11,
                          W_BITS)
            dix, diy = dt[..., 0].astype(np.int64), dt[..., 1].astype(np.int64)
            a11 = f32(_tensor_sum_f32(dix * dix) * FLT_SCALE)
            a12 = f32(_tensor_sum_f32(dix * diy) * FLT_SCALE)
            a22 = f32(_tensor_sum_f32(diy * diy) * FLT_SCALE)
            d = f32(a11 * a22 - a12 * a12)
            min_eig = f32((a22 + a11 - np.sqrt(f32((a11 - a22) * (a11 - a22) + f32(4.0) * a12 * a12))) / f32(2 * WIN * WIN))
            if min_eig < MIN_EIG or d < np.finfo(np.float32).eps:
                if level == 0:
                    status[p] = 0
                continue
            d = f32(1.0) / d
            nx, ny = nx - f32(HALF_WIN), ny - f32(HALF_WIN)
            pdx = pdy = f32(0)
            sj = pyr_j[level]
            for j in range(MAX_ITERS):
                jx, jy = int(math.floor(nx)), int(math.floor(ny))
                if jx < -WIN or jx >= cols or jy < -WIN or jy >= rows:
                    if level == 0:
                        status[p] = 0
                    break
                w00, w01, w10, w11 = _weights(nx - f32(jx), ny - f32(jy))
                y1, x1 = jy + WIN, jx + WIN
                diff = _descale(sj[y1:y1 + WIN, x1:x1 + WIN] * w00 + sj[y1:y1 + WIN, x1 + 1:x1 + WIN + 1] * w01
                                + sj[y1 + 1:y1 + WIN + 1, x1:x1 + WIN] * w10 + sj[y1 + 1:y1 + WIN + 1, x1 + 1:x1 + WIN + 1] * w11,
                                W_BITS - 5) - tmpl
                b1 = f32(_mismatch_sum_f32(diff.astype(np.int64) * dix) * FLT_SCALE)
                b2 = f32(_mismatch_sum_f32(diff.astype(np.int64) * diy) * FLT_SCALE)
                dx = f32(f32(a12 * b2 - a22 * b1) * d)
                dy = f32(f32(a12 * b1 - a11 * b2) * d)
                nx, ny = nx + dx, ny + dy
                out[p] = (nx + f32(HALF_WIN), ny + f32(HALF_WIN))
                if dx * dx + dy * dy <= EPS_SQ:
                    break
                if j > 0 and abs(dx + pdx) < 0.01 and abs(dy + pdy) < 0.01:
                    out[p, 0] -= dx * f32(0.5)
                    out[p, 1] -= dy * f32(0.5)
                    break
                pdx, pdy = dx, dy
            # the next level continues from the stored point
            nx, ny = out[p]
            if status[p] and level == 0:
                ex, ey = out[p, 0] - f32(HALF_WIN), out[p, 1] - f32(HALF_WIN)
                jx, jy = int(math.floor(ex)), int(math.floor(ey))
                if jx < -WIN or jx >= cols or jy < -WIN or jy >= rows:
                    status[p] = 0
                    continue
                w00, w01, w10, w11 = _weights(ex - f32(jx), ey - f32(jy))
                y1, x1 = jy + WIN, jx + WIN
                diff = _descale(sj[y1:y1 + WIN, x1:x1 + WIN] * w00 + sj[y1:y1 + WIN, x1 + 1:x1 + WIN + 1] * w01
                                + sj[y1 + 1:y1 + WIN + 1, x1:x1 + WIN] * w10 + sj[y1 + 1:y1 + WIN + 1, x1 + 1:x1 + WIN + 1] * w11,
                                W_BITS - 5) - tmpl
                err[p] = f32(np.abs(diff).sum()) / f32(32 * WIN * WIN)      # `errval * 1.f / (32 * w * h)`: a float division
    return out, status, err
