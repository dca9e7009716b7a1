"""Overlay drawing (reference draw.py) is visualisation and out of scope.
The mixin keeps the method names ``PoseDetector`` calls; drawing is a no-op unless OpenCV's
imgproc is importable, in which case a minimal overlay is produced."""
import numpy as np


class Draw:
    def __init__(self, logger):
        self.logger = logger

    @staticmethod
    def _cv():
        try:
            import cv2
            return cv2
        except ImportError:           # headless / not installed: drawing is optional
            return None

    @classmethod
    def draw_corners(cls, img: np.ndarray, detection) -> None:
        cv = cls._cv()
        if cv is None or img is None:
            return
        pts = np.asarray(detection.corners).astype(int).reshape(4, 2)
        cv.polylines(img, [pts.reshape(-1, 1, 2)], True, (0, 255, 0), 2)

    def draw_squares_and_3d_pts(self, img: np.ndarray, draw_frame: np.ndarray, imgpts: np.ndarray) -> None:
        cv = self._cv()
        if cv is None or img is None:
            return
        h, w = img.shape[:2]
        for p in np.asarray(imgpts).reshape(-1, 2):
            if np.all(np.isfinite(p)) and 0 <= p[0] < w and 0 <= p[1] < h:
                cv.circle(img, (int(p[0]), int(p[1])), 3, (0, 0, 255), -1)
