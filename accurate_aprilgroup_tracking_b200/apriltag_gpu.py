"""``apriltag``-shaped detector on the B200.

The reference builds ``apriltag.Detector(options)`` every frame and calls ``detector.detect(gray, return_image=True)``
(detect_pose.py:86-95, :368-371); the swatbotics library behind it is not vendored.  This module offers the surface the
reference uses - ``DetectorOptions(**kw)``, ``Detector(options).detect(img, return_image=False)``, detections with ``tag_family,
tag_id, hamming, goodness, decision_margin, homography, center, corners`` and ``tostring()`` - on top of ``agt_detect_tags``
(csrc/agt_tags.cu).  ``PoseDetector`` uses it when no ``apriltag`` module is installed.

Conventions: ``corners`` (4,2) in the order the reference's object points assume (transform_helper.py:56-59: (-,-), (-,+), (+,+),
(+,-) in tag coordinates); ``decision_margin`` = mean distance of the 64 cell means from the black/white threshold (the
reference keeps detections with ``decision_margin >= 50``, detect_pose.py:389); ``homography`` maps tag coordinates in [-1, 1]^2
(y up) to pixels.  Only the family the reference uses, tag36h11, is available.
"""
from __future__ import annotations

import numpy as np

from . import cv_compat


class DetectorOptions:
    def __init__(self, families="tag36h11", border=1, nthreads=4, quad_decimate=1.0, quad_blur=0.0, refine_edges=True,
                 refine_decode=False, refine_pose=False, debug=False, quad_contours=True):
        if families != "tag36h11":
            raise ValueError("only the tag36h11 family is available on the device")
        self.families, self.border, self.nthreads = families, int(border), int(nthreads)
        self.quad_decimate, self.quad_blur = float(quad_decimate), float(quad_blur)
        self.refine_edges, self.refine_decode, self.refine_pose = bool(refine_edges), bool(refine_decode), bool(refine_pose)
        self.debug, self.quad_contours = bool(debug), bool(quad_contours)


def _homography(corners: np.ndarray) -> np.ndarray:
    """3x3 map of tag coordinates (-1,-1), (-1,1), (1,1), (1,-1) (the corner order) onto the four corners."""
    src = np.array([[-1.0, -1.0], [-1.0, 1.0], [1.0, 1.0], [1.0, -1.0]])
    a, b = [], []
    for (x, y), (u, v) in zip(src, corners):
        a.append([x, y, 1, 0, 0, 0, -u * x, -u * y]); b.append(u)
        a.append([0, 0, 0, x, y, 1, -v * x, -v * y]); b.append(v)
    h = np.linalg.solve(np.array(a), np.array(b))
    return np.append(h, 1.0).reshape(3, 3)


class Detection:
    def __init__(self, tag_id, corners, decision_margin, hamming):
        self.tag_family = b"tag36h11"
        self.tag_id = int(tag_id)
        self.hamming = int(hamming)
        self.goodness = 0.0
        self.decision_margin = float(decision_margin)
        self.corners = np.asarray(corners, dtype=np.float64).reshape(4, 2)
        self.homography = _homography(self.corners)
        c = self.homography @ np.array([0.0, 0.0, 1.0])
        self.center = c[:2] / c[2]

    def tostring(self, values=None, indent=0):
        return " " * indent + "Detection(tag_family=%r, tag_id=%d, hamming=%d, decision_margin=%.2f)" % (
            self.tag_family, self.tag_id, self.hamming, self.decision_margin)

    __str__ = tostring


class Detector:
    def __init__(self, options=None, searchpath=None, context=None):
        self.options = options if options is not None else DetectorOptions()
        self._ctx = context if context is not None else cv_compat.default_context()

    def detect(self, img, return_image=False):
        """-> list of Detection sorted by tag id (and a zero image of the frame's size when ``return_image``: the reference ignores
        it, detect_pose.py:371)."""
        gray = np.asarray(img)
        if gray.ndim != 2 or gray.dtype != np.uint8:
            raise ValueError("detect() takes a single-channel uint8 image")
        win = 4 if self.options.refine_edges else 0
        dets = [Detection(t, c, m, hd) for t, c, m, hd in self._ctx.detect_tags(gray, refine_win=win)]
        if return_image:
            return dets, np.zeros(gray.shape, dtype=np.uint8)
        return dets
