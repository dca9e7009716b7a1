"""CPU oracle of the whole hot path for one camera stream: APE -> (LK when < 2 tags) -> dense refinement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Composition of the three stage oracles in the order the
path runs them: the reference's APE state machine (oracle/ape_oracle.py, pinned to the
unmodified reference), cv2.calcOpticalFlowPyrLK for the inter-frame corner tracking that backs up frames
with fewer than two detected tags (SURVEY.md 9.2 integration rule: a tag is re-admitted when all four of
its corners have status 1), and oracle/dpr_oracle.py for the refinement that replaces the accepted PnP
pose before it becomes prev_transform (detect_pose.py:569).  The mounted reference contains only the first
stage, so the composition itself is frozen here; BASELINE config 1 (640x480 synthetic sequence) runs it.
"""
from __future__ import annotations

import copy
from typing import List, Optional, Sequence, Tuple

import cv2 as cv
import numpy as np

from oracle import ape_oracle, dpr_oracle, lk_oracle


class PipelineOracle(ape_oracle.ApeOracle):
    def __init__(self, group, mtx, model: dpr_oracle.Model, use_lk=True, use_dense_refine=True, dist=None):
        """``dist``: the lens model the reference hands to solvePnP / projectPoints (detect_pose.py:509-526) - also for frames
        that process_frame has already undistorted, a quirk of the reference that is kept.  Dense refinement reads pixels and
        uses ``refine_mtx`` of the frame (see ``frame``), never ``dist``."""
        super().__init__(group, mtx, dist, True)
        self.model = model
        self.use_lk, self.use_dense_refine = use_lk, use_dense_refine
        self.prev_gray: Optional[np.ndarray] = None
        self.prev_corners: List[Tuple[int, np.ndarray]] = []
        self.gray: Optional[np.ndarray] = None
        self.tracked = 0

    def _refine(self, pose):
        """Dense refinement in place (the result lives in the same arrays solvePnP returned, dtype preserved)."""
        init = np.concatenate([np.asarray(pose[0], dtype=np.float64).ravel(), np.asarray(pose[1], dtype=np.float64).ravel()])
        out = dpr_oracle.refine(lk_oracle.pyramid_cv(self.gray, 4), self.model, self.refine_mtx, init)
        if out["status"] != dpr_oracle.ST_NONE:
            pose[0].reshape(-1)[:] = out["pose"][:3]
            pose[1].reshape(-1)[:] = out["pose"][3:]

    def frame(self, gray: np.ndarray, dets: Sequence[Tuple[int, np.ndarray]], refine_mtx=None):
        """One frame: gray (H,W) u8 + accepted detections [(tag_id, corners (4,2))].  ``refine_mtx``: pinhole camera matrix of the
        pixels in ``gray`` when it is not ``mtx`` - for a frame undistorted and cropped by undistort_frame (detect_pose.py:147-183)
        the new camera matrix with its principal point moved by the crop offset."""
        self.prev_gray, self.gray = self.gray, gray
        self.refine_mtx = self.mtx if refine_mtx is None else np.asarray(refine_mtx, dtype=np.float64)
        dets = [(t, np.asarray(c, dtype=np.float64).reshape(4, 2)) for t, c in dets]
        self.tracked = 0
        if self.use_lk and len(dets) < ape_oracle.MIN_TAGS and self.prev_gray is not None and self.prev_corners:
            have = {t for t, _ in dets}
            todo = [(t, c) for t, c in self.prev_corners if t not in have]
            if todo:
                pts = np.concatenate([c for _, c in todo]).astype(np.float32)
                nxt, st, _ = lk_oracle.lk_cv(self.prev_gray, self.gray, pts)
                for k, (t, _) in enumerate(todo):
                    if st[4 * k:4 * k + 4].all():
                        dets.append((t, nxt[4 * k:4 * k + 4].astype(np.float64)))
                        self.tracked += 1
        self._dense_hook = self.use_dense_refine
        self.step(dets)
        self.prev_corners = dets if self.last_accepted else []

    # the APE step with the refinement hook between the gate and the state update (detect_pose.py:539-569)
    def step(self, dets):
        held_prev = copy.deepcopy(self.prev)
        self.last_error, self.last_accepted = None, False
        if len(dets) < ape_oracle.MIN_TAGS:
            self.guess = (None, None)
            return
        obj = np.array([ape_oracle.tag_object_points(*self._sz_r_t(t)) for t, _ in dets], dtype=np.float32).reshape(-1, 3)
        img = np.array([np.asarray(c).reshape(1, 4, 2) for _, c in dets], dtype=np.float32).reshape(-1, 2)
        fresh = self.guess[0] is None or not self.enhance_ape
        if fresh:
            ok, rvec, tvec = cv.solvePnP(obj, img, self.mtx, self.dist, flags=cv.SOLVEPNP_ITERATIVE)
        else:
            ok, rvec, tvec = cv.solvePnP(obj, img, self.mtx, self.dist, self.guess[0], self.guess[1], True,
                                         flags=cv.SOLVEPNP_ITERATIVE)
        pose = (rvec, tvec)
        if not ok:
            return
        err = ape_oracle.mean_reprojection_error(obj, img, rvec, tvec, self.mtx, self.dist)
        self.last_error = float(err)
        if not err < ape_oracle.MAX_MEAN_ERROR:
            self.guess = (None, None)
            return
        self.last_accepted = True
        if getattr(self, "_dense_hook", False) and self.gray is not None:
            self._refine(pose)
        if fresh:
            self.guess = pose
        else:
            good, tran, rot, tacc, racc = self._velocities(pose, held_prev)
            if good:
                self.guess = self._predict(held_prev, tran, tacc, rot, racc)
        self.prev = pose
