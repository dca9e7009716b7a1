"""Minimal stand-in for the swatbotics ``apriltag`` module (absent in this image).

TEST INFRASTRUCTURE: lets the *unmodified* reference import
(detect_pose.py:22, draw.py:7) and feeds it precomputed synthetic detections.
Surface used by the reference: DetectorOptions(**kw) (detect_pose.py:86-95),
Detector(options).detect(img, return_image=True) (detect_pose.py:368-371),
Detection fields + tostring() (detect_pose.py:389-400, draw.py:156-192).
"""
import collections

import numpy as np

_QUEUE = collections.deque()


def push_detections(dets):
    """Queue the detections the next ``Detector.detect`` call returns."""
    _QUEUE.append(list(dets))


class DetectorOptions:
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)


class Detection:
    def __init__(self, tag_id, corners, decision_margin=100.0):
        self.tag_family = b"tag36h11"
        self.tag_id = int(tag_id)
        self.hamming = 0
        self.goodness = 0.0
        self.decision_margin = float(decision_margin)
        self.corners = np.asarray(corners, dtype=np.float64).reshape(4, 2)
        self.center = self.corners.mean(axis=0)
        self.homography = np.eye(3)

    def tostring(self, values=None, indent=0):
        return " " * indent + "Detection(tag_id=%d, margin=%.1f)" % (self.tag_id, self.decision_margin)


class Detector:
    def __init__(self, options=None, searchpath=None):
        self.options = options

    def detect(self, img, return_image=False):
        dets = _QUEUE.popleft() if _QUEUE else []
        if return_image:
            return dets, np.zeros(img.shape[:2], dtype=np.uint8)
        return dets
