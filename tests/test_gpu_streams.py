"""GPU test of the batched multi-stream pipeline (BASELINE config 5 shape) against the single-stream
drop-in ``PoseDetector`` driven frame by frame with the same frames and detections."""
import logging

import numpy as np
import pytest

from accurate_aprilgroup_tracking_b200 import synth
from tests import util
from tests.test_gpu_dropin import _Det, _logger, detector_factory  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def test_batched_streams_match_single_stream_detector(ctxvga, detector_factory):
    from accurate_aprilgroup_tracking_b200.batched import BatchedPoseDetector, pack_detections
    cam = synth.CAMERA_VGA
    n_streams, n_frames = 4, 14
    trajs = [synth.trajectory(8000 + s, n_frames) for s in range(n_streams)]
    rngs = [np.random.default_rng(8000 + s) for s in range(n_streams)]
    bpd = BatchedPoseDetector(ctxvga, n_streams, cam.width, cam.height, synth.object_points())
    singles = [detector_factory(cam.mtx, use_lk=True, use_dense_refine=True) for _ in range(n_streams)]
    tracked_total = 0
    for f in range(n_frames):
        poses = np.array([trajs[s][f] for s in range(n_streams)])
        ctxvga.render(bpd.pyr[bpd.cur], poses, np.arange(n_streams) + 100 * f)
        frames = bpd.frames.cpu().numpy().copy()
        dets = []
        for s in range(n_streams):
            d = synth.detections(trajs[s][f], cam, rngs[s])
            if s == 1 and f in (5, 6):
                d = d[:1]                       # detector drops to one tag: LK must carry the others
            if s == 2 and f == 8:
                d = []                          # nothing detected at all
            dets.append(d)
        img, valid, ntags = pack_detections(dets)
        out = bpd.step(img, valid, ntags)
        pose_b = out["pose"].cpu().numpy()
        acc_b = out["accepted"].cpu().numpy()
        if out["tracked_tags"] is not None:
            tracked_total += int(out["tracked_tags"].sum())
        for s in range(n_streams):
            det = singles[s]
            det.img = None
            det._set_gray(frames[s])
            lists = det._lists_from_detections([_Det(t, c) for t, c in dets[s]])
            if len(lists[0]) < 2:
                lists = det._track_lost_tags(*lists)
            before = None if det.prev_transform[0] is None else det.prev_transform[0].copy()
            det._estimate_pose(lists[0], lists[1])
            accepted_single = bool(det._prev_corners)
            assert bool(acc_b[s]) == accepted_single, (f, s)
            if det.prev_transform[0] is not None:
                got = np.concatenate([det.prev_transform[0].ravel(), det.prev_transform[1].ravel().astype(np.float64)])
                util.assert_pose_close(pose_b[s], got, f"frame {f} stream {s}")
            if accepted_single:
                dr, dt = util.pose_diff(pose_b[s], trajs[s][f])
                assert dr < 0.02 and dt < 2e-3
    assert tracked_total >= 2          # the LK path was exercised
