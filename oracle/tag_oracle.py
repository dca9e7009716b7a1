"""CPU oracle for tag identification and detection.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference's detector is the un-vendored swatbotics `apriltag` C library
(detect_pose.py:86-95 tag36h11, :368-371 detect, :389-400 the `decision_margin < 50` filter and the corner order of
transform_helper.py:56-59); it is absent from the image.  What the image has is OpenCV's ArUco module with the same family
(cv2.aruco.DICT_APRILTAG_36h11, the table synth.TAG36H11_CODES was dumped from), so:

  `detect_cv`     cv2.aruco.ArucoDetector.detectMarkers - the oracle for WHICH tags are in a frame and where (its corners are
                  contour points, ~1 px from the true corners); ids and corner order converted to the reference's conventions
  `decode_np`     the frozen specification of the GPU identification step (csrc/agt_tags.cu): 8x8 cells (6x6 data + black border)
                  sampled through the quad's homography, threshold half way between the darkest and the brightest cell,
                  border check, dictionary match over the four rotations.  Pinned by the tests: on every tag cv2.aruco
                  finds, decode_np(aruco's corners) returns aruco's id; on the true corners it returns the true id, rotation 0.

Corner order everywhere: the reference's (-,-), (-,+), (+,+), (+,-) in tag coordinates with y up, i.e. bottom-left, top-left,
top-right, bottom-right of the tag as printed (transform_helper.py:56-59).
"""
from __future__ import annotations

import cv2 as cv
import numpy as np

from accurate_aprilgroup_tracking_b200 import synth

CELLS = 8                 # 6x6 data cells + one cell of black border on each side
SUB = (-0.25, 0.0, 0.25)  # sample offsets inside a cell, in cells
MAX_BORDER_ERRORS = 2
MAX_HAMMING = 2


def code_bits(code: int) -> np.ndarray:
    """(6,6) bits of a tag36h11 code word, row 0 = top row of the tag (synth.tag_cells)."""
    return np.array([(code >> (35 - i)) & 1 for i in range(36)], np.uint8).reshape(6, 6)


def detect_cv(gray: np.ndarray):
    """-> list of (tag_id, corners (4,2) f64 in the reference's order) from cv2.aruco (corners there: TL, TR, BR, BL)."""
    det = cv.aruco.ArucoDetector(cv.aruco.getPredefinedDictionary(cv.aruco.DICT_APRILTAG_36h11), cv.aruco.DetectorParameters())
    corners, ids, _ = det.detectMarkers(gray)
    out = []
    if ids is not None:
        for c, i in zip(corners, ids.ravel().tolist()):
            c = c.reshape(4, 2).astype(np.float64)
            out.append((int(i), np.stack([c[3], c[0], c[1], c[2]])))
    return out


def square_to_quad(q: np.ndarray):
    """Projective map of the unit square (s right, t down; (0,0) = top-left) onto the quad with corners in the reference's order
    (BL, TL, TR, BR) -> coefficients (a, b, c, d, e, f, g, h) of x = (a s + b t + c) / (g s + h t + 1), y = (d s + e t + f) / (...)."""
    (x3, y3), (x0, y0), (x1, y1), (x2, y2) = [tuple(map(float, p)) for p in q]       # TL=(0,0) TR=(1,0) BR=(1,1) BL=(0,1)
    dx1, dx2, sx = x1 - x2, x3 - x2, x0 - x1 + x2 - x3
    dy1, dy2, sy = y1 - y2, y3 - y2, y0 - y1 + y2 - y3
    den = dx1 * dy2 - dx2 * dy1
    g = (sx * dy2 - dx2 * sy) / den
    h = (dx1 * sy - sx * dy1) / den
    return (x1 - x0 + g * x1, x3 - x0 + h * x3, x0, y1 - y0 + g * y1, y3 - y0 + h * y3, y0, g, h)


def cell_means(gray: np.ndarray, quad: np.ndarray) -> np.ndarray:
    """(8,8) mean intensity of 3x3 bilinear samples per cell; NaN if a sample leaves the image."""
    a, b, c, d, e, f, g, h = square_to_quad(quad)
    hh, ww = gray.shape
    out = np.zeros((CELLS, CELLS))
    img = gray.astype(np.float64)
    for r in range(CELLS):
        for cc in range(CELLS):
            acc = 0.0
            for dt in SUB:
                for ds in SUB:
                    s, t = (cc + 0.5 + ds) / CELLS, (r + 0.5 + dt) / CELLS
                    w = g * s + h * t + 1.0
                    x, y = (a * s + b * t + c) / w, (d * s + e * t + f) / w
                    x0, y0 = int(np.floor(x)), int(np.floor(y))
                    if x0 < 0 or y0 < 0 or x0 + 1 >= ww or y0 + 1 >= hh:
                        return np.full((CELLS, CELLS), np.nan)
                    fx, fy = x - x0, y - y0
                    acc += ((1 - fx) * (1 - fy) * img[y0, x0] + fx * (1 - fy) * img[y0, x0 + 1]
                            + (1 - fx) * fy * img[y0 + 1, x0] + fx * fy * img[y0 + 1, x0 + 1])
            out[r, cc] = acc / 9.0
    return out


def decode_np(gray: np.ndarray, quad: np.ndarray, codes=None, max_hamming: int = MAX_HAMMING):
    """-> (tag_id or -1, rotation 0..3, hamming, margin).  rotation k: the quad's first corner is the tag's corner k (so the
    tag's corners in the reference's order are np.roll(quad, -k, axis=0)); margin = mean |cell mean - threshold| (the
    analogue of apriltag's decision_margin, which the reference compares with 50, detect_pose.py:389)."""
    codes = synth.TAG36H11_CODES if codes is None else codes
    m = cell_means(gray, np.asarray(quad, dtype=np.float64))
    if np.isnan(m).any():
        return -1, 0, 255, 0.0
    thr = 0.5 * (m.min() + m.max())
    bits = (m > thr).astype(np.uint8)
    margin = float(np.abs(m - thr).mean())
    border = np.concatenate([bits[0], bits[-1], bits[1:-1, 0], bits[1:-1, -1]])
    if int(border.sum()) > MAX_BORDER_ERRORS:
        return -1, 0, 255, margin
    data = bits[1:-1, 1:-1]
    best = (-1, 0, 255)
    for rot in range(4):
        # the quad starts at tag corner `rot`: its sampled grid is the tag's grid turned by rot quarter turns
        grid = np.rot90(data, -rot)
        word = 0
        for v in grid.ravel():
            word = (word << 1) | int(v)
        for tag_id, code in enumerate(codes):
            hd = bin(word ^ int(code)).count("1")
            if hd < best[2]:
                best = (tag_id, rot, hd)
    if best[2] > max_hamming:
        return -1, 0, best[2], margin
    return best[0], best[1], best[2], margin
