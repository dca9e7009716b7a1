"""``BatchedPoseDetector``: the full hot path (APE -> LK -> dense refinement) for many camera
streams at once, one frame per stream per step, entirely on one GPU.

It is the batched counterpart of ``PoseDetector._detect_and_get_pose`` (detect_pose.py:576-609) with
the reference's per-stream state (detect_pose.py:74-78) kept in device memory.  Per step:

  K1  agt_build_pyramid    pyramid of the new frames
  K2  agt_lk + agt_lk_merge  streams with < 2 detected tags re-admit the tags whose four corners
                           were tracked from the previous frame (the inlier set is LK status == 1)
  K0  agt_ape_prepare      extrinsic guess from the predictor state        (detect_pose.py:508)
  K3  agt_pnp              batched solvePnP + mean reprojection error      (detect_pose.py:509-538)
  K4  agt_refine           dense refinement of the poses that pass the 2 px gate (masked)
  K0  agt_ape_update       accept / reset rules, velocity FIFOs, predictor (detect_pose.py:539-574)

Frames within a stream are sequential (the predictor needs poses t-1, t-2; LK needs frame t-1), so the
batch dimension is the number of streams; streams are pinned to GPUs by ``sharding.stream_owner``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import synth
from .context import AgtContext


class BatchedPoseDetector:
    def __init__(self, ctx: AgtContext, n_streams: int, width: int, height: int, obj_pts: np.ndarray,
                 enhance_ape: bool = True, use_lk: bool = True, use_dense_refine: bool = True, levels: int = 4):
        t = ctx.torch
        self.ctx, self.n, self.enhance_ape = ctx, int(n_streams), enhance_ape
        self.use_lk, self.use_dense_refine = use_lk, use_dense_refine
        self.n_pts = int(obj_pts.shape[0])
        self.obj = ctx._dev(obj_pts, t.float32)
        self.pyr = [ctx.alloc_pyramid(self.n, width, height, levels) for _ in range(2)]
        self.cur = 0
        self.state = ctx.new_stream_state(self.n)
        dev = ctx.tdev
        self.prev_pts = t.zeros((self.n, self.n_pts, 2), dtype=t.float32, device=dev)
        self.prev_valid = t.zeros((self.n, self.n_pts), dtype=t.uint8, device=dev)
        self.have_prev_frame = False
        if use_dense_refine and ctx._model is None:
            ctx.set_synthetic_model()

    @property
    def frames(self):
        """Level-0 buffer [S,H,W] the caller fills (or renders into) before ``step``."""
        return self.pyr[self.cur].frames

    def step(self, img_pts, valid, n_tags, frames=None):
        """img_pts [S,P,2] f32, valid [S,P] u8 (corner-level; all four corners of a detected tag set),
        n_tags [S] i32 accepted detections.  ``frames`` [S,H,W] u8 is copied into the current slot
        unless the caller wrote ``self.frames`` directly.  Returns a dict of device tensors."""
        ctx, t = self.ctx, self.ctx.torch
        cur, prv = self.pyr[self.cur], self.pyr[1 - self.cur]
        if frames is not None:
            ctx.upload_frames(cur, frames)
        img = ctx._dev(img_pts, t.float32).clone()
        val = ctx._dev(valid, t.uint8).clone()
        ntg = ctx._dev(n_tags, t.int32).clone()
        ctx.build_pyramid(cur)                                                   # K1
        tracked_tags = None
        if self.use_lk and self.have_prev_frame:
            nxt, st, _ = ctx.lk(prv, cur, self.prev_pts, n_tags=ntg)              # K2 (frames with < 2 tags only)
            before = ntg.clone()
            ctx.lk_merge(nxt, st, self.prev_valid, img, val, ntg)
            tracked_tags = ntg - before
        guess, use = ctx.ape_prepare(self.state, self.enhance_ape)               # K0
        pose, ok, err, iters = ctx.pnp(self.obj, img, val, guess, use)            # K3
        refined = None
        if self.use_dense_refine:
            gate = ((ok != 0) & (err < 2.0) & (ntg >= 2)).to(t.uint8)             # only poses the reference accepts
            refined = ctx.refine(cur, pose.reshape(self.n, 1, 6), 1, mask=gate)   # K4
            good = (refined["status"].reshape(self.n) != 0).unsqueeze(1)
            pose = t.where(good, refined["pose"].reshape(self.n, 6), pose)
        accepted, flag = ctx.ape_update(self.state, ntg, pose, ok, err, self.enhance_ape)   # K0
        # corners of the accepted frame feed the next LK step (PoseDetector._prev_corners)
        self.prev_pts = img
        self.prev_valid = val * accepted.unsqueeze(1)
        self.have_prev_frame = True
        self.cur = 1 - self.cur
        return {"pose": self.state[:, 1:7].clone(), "accepted": accepted, "error_flag": flag, "reproj_err": err,
                "n_tags": ntg, "tracked_tags": tracked_tags, "refine": refined}


def pack_detections(dets_per_stream, n_tags_total: int = synth.NUM_TAGS):
    """[(tag_id, corners (4,2)), ...] per stream -> (img_pts [S,4T,2] f32, valid [S,4T] u8, n_tags [S] i32)."""
    s = len(dets_per_stream)
    img = np.zeros((s, 4 * n_tags_total, 2), np.float32)
    valid = np.zeros((s, 4 * n_tags_total), np.uint8)
    n = np.zeros(s, np.int32)
    for i, dets in enumerate(dets_per_stream):
        n[i] = len(dets)
        for tag, corners in dets:
            img[i, 4 * tag:4 * tag + 4] = corners
            valid[i, 4 * tag:4 * tag + 4] = 1
    return img, valid, n
