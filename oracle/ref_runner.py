"""Drive the UNMODIFIED reference (/root/reference) on synthetic detections.

TEST INFRASTRUCTURE.  Only usable in the build container: /root/reference does
not exist on the GPU box, so this module is used (a) by ``make_golden.py`` to
generate the committed fixtures and (b) by CPU tests that skip when the mount
is absent.  It imports ``aprilgroup_pose_estimation.detect_pose.PoseDetector``
exactly as main.py:8-10 does, with ``oracle/apriltag_stub`` standing in for the
missing ``apriltag`` module, and runs ``_detect_and_get_pose``
(detect_pose.py:576-609) / ``_estimate_pose`` (detect_pose.py:467-574).
"""
from __future__ import annotations

import contextlib
import io
import logging
import os
import sys
import tempfile
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

REFERENCE_ROOT = Path("/root/reference/aprilgroup_tracking")
_STUB_DIR = str(Path(__file__).resolve().parent / "apriltag_stub")


def reference_available() -> bool:
    return (REFERENCE_ROOT / "aprilgroup_pose_estimation" / "detect_pose.py").exists()


def _null_logger() -> logging.Logger:
    lg = logging.getLogger("agt-oracle-null")
    lg.handlers[:] = [logging.NullHandler()]
    lg.propagate = False
    lg.setLevel(logging.CRITICAL)
    return lg


class ReferenceRunner:
    """Owns one reference ``PoseDetector`` instance built over the synthetic group."""

    def __init__(self, mtx: np.ndarray, dist: Optional[np.ndarray], enhance_ape: bool = True,
                 workdir: Optional[str] = None):
        if not reference_available():
            raise RuntimeError("reference sources are not mounted at /root/reference")
        from accurate_aprilgroup_tracking_b200 import synth
        for p in (_STUB_DIR, str(REFERENCE_ROOT)):
            if p not in sys.path:
                sys.path.insert(0, p)
        self._tmp = tempfile.TemporaryDirectory() if workdir is None else None
        self.workdir = workdir or self._tmp.name
        synth.write_april_group_json(Path(self.workdir))
        import apriltag as stub                                   # noqa: the stub
        from aprilgroup_pose_estimation.detect_pose import PoseDetector  # the reference
        self.stub = stub
        cwd = os.getcwd()
        os.chdir(self.workdir)                                    # DIRPATH is CWD-relative (detect_pose.py:54)
        try:
            self.det = PoseDetector(_null_logger(), mtx, dist, enhance_ape)
        finally:
            os.chdir(cwd)

    # -- per-frame drivers ----------------------------------------------------
    def estimate(self, dets: Sequence[Tuple[int, np.ndarray]]):
        """One frame through ``_obtain_detections``-equivalent list building and
        the reference ``_estimate_pose``.  Returns a snapshot of the state."""
        img_list, obj_list = [], []
        for tag_id, corners in dets:
            size, tvec, rvec = self.det.extrinsics[tag_id]
            obj = self.det.transform_marker_corners(self.det.get_initial_pts(size), (rvec, tvec))
            img_list.append(np.asarray(corners).reshape(1, 4, 2))
            obj_list.append(obj)
        self.det.img = np.zeros((8, 8, 3), np.uint8)
        self.det.draw_frame = np.zeros((8, 8, 3), np.uint8)
        with contextlib.redirect_stdout(io.StringIO()):           # transform_helper.py:218 prints
            self.det._estimate_pose(img_list, obj_list)
        return self.snapshot()

    def detect_and_get_pose(self, frame_bgr: np.ndarray, dets: Sequence[Tuple[int, np.ndarray]],
                            margins: Optional[Sequence[float]] = None):
        """Full reference per-frame entry point with stubbed detector output."""
        margins = margins if margins is not None else [100.0] * len(dets)
        self.stub.push_detections(self.stub.Detection(t, c, m) for (t, c), m in zip(dets, margins))
        with contextlib.redirect_stdout(io.StringIO()):
            self.det._detect_and_get_pose(frame_bgr)
        return self.snapshot()

    def snapshot(self):
        def cp(t):
            return None if t[0] is None else (np.array(t[0], dtype=np.float64).reshape(3),
                                              np.array(t[1], dtype=np.float64).reshape(3))
        return {"prev": cp(self.det.prev_transform), "guess": cp(self.det.extrinsic_guess),
                "n_vel": len(self.det.rot_velocities)}
