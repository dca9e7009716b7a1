"""CPU oracle for sub-pixel corner refinement (first step: refinement around predicted corners).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference's tag detector is the un-vendored swatbotics `apriltag` C
library (detect_pose.py:86-95, :368-371, absent from the image), whose `refine_edges` option snaps the quad corners to the
image gradient.  What the image does have is OpenCV, whose ArUco module detects the same tag36h11 family and refines
corners with cv2.cornerSubPix; the frozen semantics of the GPU corner refinement are therefore

    cv2.cornerSubPix(gray, corners, (win, win), (-1, -1), (COUNT + EPS, max_iters, eps))

`corner_subpix_cv` calls it (the oracle proper); `corner_subpix_np` restates the published algorithm (OpenCV
modules/imgproc/src/cornersubpix.cpp + getRectSubPix: Gaussian-weighted gradient normal equations over a (2 win + 1)^2
window of a bilinearly resampled float patch, replicated border, central differences, 2x2 solve in float64, the point kept if
it moves further than the window) and is pinned to it by tests/test_cpu_oracle.py.
"""
from __future__ import annotations

import cv2 as cv
import numpy as np

WIN = 5
MAX_ITERS = 30
EPS = 1e-3


def corner_subpix_cv(gray: np.ndarray, pts: np.ndarray, win: int = WIN, max_iters: int = MAX_ITERS, eps: float = EPS) -> np.ndarray:
    p = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 1, 2).copy()
    cv.cornerSubPix(gray, p, (win, win), (-1, -1), (cv.TERM_CRITERIA_COUNT + cv.TERM_CRITERIA_EPS, max_iters, eps))
    return p.reshape(-1, 2)


def _patch(gray: np.ndarray, cx: np.float32, cy: np.float32, size: int) -> np.ndarray:
    """cv::getRectSubPix(gray, (size, size), (cx, cy)) as float32: bilinear, BORDER_REPLICATE, the x fraction at least 1e-4."""
    f32 = np.float32
    h, w = gray.shape
    x0 = f32(cx - f32((size - 1) * 0.5))
    y0 = f32(cy - f32((size - 1) * 0.5))
    ix, iy = int(np.floor(x0)), int(np.floor(y0))
    a = max(f32(x0 - f32(ix)), f32(0.0001))
    b = f32(y0 - f32(iy))
    xs = np.clip(ix + np.arange(size + 1), 0, w - 1)
    ys = np.clip(iy + np.arange(size + 1), 0, h - 1)
    g = gray[np.ix_(ys, xs)].astype(np.float32)
    top = g[:-1] * f32(f32(1) - b) + g[1:] * b                    # rows blended first (b1 p + b2 p_below), float32
    return (top[:, :-1] * f32(f32(1) - a) + top[:, 1:] * a).astype(np.float32)


def corner_subpix_np(gray: np.ndarray, pts: np.ndarray, win: int = WIN, max_iters: int = MAX_ITERS, eps: float = EPS) -> np.ndarray:
    f32 = np.float32
    h, w = gray.shape
    n = 2 * win + 1
    k = (np.arange(n, dtype=np.float32) - f32(win)) / f32(win)
    g1 = np.exp(-k * k).astype(np.float32)
    mask = (g1[:, None] * g1[None, :]).astype(np.float32).astype(np.float64)
    px = np.arange(n, dtype=np.float64)[None, :] - win
    py = np.arange(n, dtype=np.float64)[:, None] - win
    out = np.array(pts, dtype=np.float32).reshape(-1, 2).copy()
    for i in range(len(out)):
        tx, ty = out[i]
        cx, cy = tx, ty
        for _ in range(max(1, min(max_iters, 100))):
            sp = _patch(gray, cx, cy, n + 2).astype(np.float64)
            tgx = sp[1:-1, 2:] - sp[1:-1, :-2]
            tgy = sp[2:, 1:-1] - sp[:-2, 1:-1]
            gxx, gxy, gyy = tgx * tgx * mask, tgx * tgy * mask, tgy * tgy * mask
            a, b, c = gxx.sum(), gxy.sum(), gyy.sum()
            bb1 = (gxx * px + gxy * py).sum()
            bb2 = (gxy * px + gyy * py).sum()
            det = a * c - b * b
            if abs(det) <= np.finfo(np.float64).eps ** 2:
                break
            s = 1.0 / det
            nx = f32(cx + c * s * bb1 - b * s * bb2)
            ny = f32(cy - b * s * bb1 + a * s * bb2)
            err = float(nx - cx) ** 2 + float(ny - cy) ** 2
            cx, cy = nx, ny
            if cx < 0 or cx >= w or cy < 0 or cy >= h:
                break
            if not err > eps * eps:
                break
        if abs(cx - tx) > win or abs(cy - ty) > win:
            cx, cy = tx, ty
        out[i] = (cx, cy)
    return out
