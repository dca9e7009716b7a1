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
        # (every other frame through the packed form of the inputs: one copy instead of three, same buffers)
        out = bpd.step(bpd.pack_inputs(img, valid, ntags)) if f % 2 else bpd.step(img, valid, ntags)
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


def _oracle_stream(job):
    """One stream through oracle/pipeline_oracle.py (a worker process): -> per frame (accepted, pose or None, tags tracked).
    The workers are SPAWNED (a fork of a process that holds a CUDA context and the BLAS / OpenCV thread pools can deadlock in the
    child) and map the frames from a file instead of receiving 2 GB through a pipe."""
    import cv2
    from oracle import ape_oracle, pipeline_oracle
    cv2.setNumThreads(1)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)                      # one BLAS thread per worker: 16 workers do not fight over 16 x 16 threads
    except Exception:
        pass
    path, stream, dets, mtx = job
    frames = np.load(path, mmap_mode="r")[:, stream]
    po = pipeline_oracle.PipelineOracle(ape_oracle.group_from_json(synth.april_group_dict()), mtx, util.dpr_model())
    out = []
    for f in range(len(frames)):
        po.frame(np.ascontiguousarray(frames[f]), dets[f])
        pose = None if po.prev[0] is None else np.concatenate([po.prev[0].ravel(), po.prev[1].ravel().astype(np.float64)])
        out.append((po.last_accepted, pose, po.tracked))
    return out


def test_config5_64_streams_1080p_match_pipeline_oracle(ctx1080, tmp_path):
    """BASELINE config 5 at its shape: 64 concurrent 1080p streams, 16 frames each, full APE + LK + dense refinement per frame
    with detector dropouts (one tag left -> LK carries the rest; nothing detected -> the stream resets), the batched detector
    against the CPU composition of the stage oracles run per stream (reference state machine + cv2 LK + dense oracle)."""
    import multiprocessing as mp
    import os
    from accurate_aprilgroup_tracking_b200.batched import BatchedPoseDetector
    cam = synth.CAMERA_1080P
    n_streams, n_frames = 64, 16
    trajs = [synth.trajectory(9000 + s, n_frames) for s in range(n_streams)]
    rngs = [np.random.default_rng(9000 + s) for s in range(n_streams)]
    bpd = BatchedPoseDetector(ctx1080, n_streams, cam.width, cam.height, synth.object_points())
    frames_all = np.empty((n_frames, n_streams, cam.height, cam.width), np.uint8)
    dets_all = []
    got_pose, got_acc, got_tracked = [], [], []
    for f in range(n_frames):
        ctx1080.render(bpd.pyr[bpd.cur], np.array([trajs[s][f] for s in range(n_streams)]), np.arange(n_streams) + 1000 * f)
        frames_all[f] = bpd.frames.cpu().numpy()
        dets = []
        for s in range(n_streams):
            d = synth.detections(trajs[s][f], cam, rngs[s])
            if (f + 3 * s) % 11 == 10:
                d = d[:1]                       # one tag left: LK has to carry the others
            if s % 16 == 5 and f == 7:
                d = []                          # nothing detected: the stream loses its guess
            dets.append(d)
        dets_all.append(dets)
        out = bpd.step(*bpd.pack(dets))
        got_pose.append(out["pose"].cpu().numpy().copy())
        got_acc.append(out["accepted"].cpu().numpy().copy())
        got_tracked.append(out["tracked_tags"].cpu().numpy().copy())
    path = str(tmp_path / "config5_frames.npy")
    np.save(path, frames_all)
    del frames_all
    jobs = [(path, s, [dets_all[f][s] for f in range(n_frames)], cam.mtx) for s in range(n_streams)]
    with mp.get_context("spawn").Pool(min(16, os.cpu_count() or 1)) as pool:
        want = pool.map_async(_oracle_stream, jobs, chunksize=1).get(timeout=1200)      # a stuck worker fails the test, not the run
    os.remove(path)
    n_checked = n_tracked = 0
    for s in range(n_streams):
        for f in range(n_frames):
            acc, pose, tracked = want[s][f]
            assert bool(got_acc[f][s]) == acc, (s, f)
            assert int(got_tracked[f][s]) == tracked, (s, f)          # LK inlier set (all-four-corners rule)
            n_tracked += tracked
            if acc:
                util.assert_pose_close(got_pose[f][s], pose, f"stream {s} frame {f}")
                n_checked += 1
    print(f"config 5: {n_checked} accepted poses of {n_streams * n_frames} checked against the pipeline oracle, {n_tracked} tags re-admitted by LK")
    assert n_checked > 0.9 * n_streams * n_frames and n_tracked >= 50


@pytest.mark.parametrize("from_pixels", [False, True])
def test_stream_groups_equal_the_single_batch(ctxvga, from_pixels):
    """batched.StreamGroups (the streams of a GPU as independent batches in flight on their own CUDA streams, frame ingest of the
    next step on side streams) against one BatchedPoseDetector over all streams: same frames, same detections with dropouts
    (or none at all: the detector on the device) -> every pose of every frame bit-identical, same accept decisions; graphs are
    captured on the way (more than two steps)."""
    import torch
    from accurate_aprilgroup_tracking_b200.batched import BatchedPoseDetector, StreamGroups, pack_detections
    from accurate_aprilgroup_tracking_b200.context import AgtContext
    cam = synth.CAMERA_VGA
    n_streams, n_frames, groups = 10, 9, 3
    trajs = [synth.trajectory(8100 + s, n_frames) for s in range(n_streams)]
    rngs = [np.random.default_rng(8100 + s) for s in range(n_streams)]
    bank = ctxvga.alloc_pyramid(n_streams * n_frames, cam.width, cam.height, 1)
    for f in range(n_frames):
        ctxvga.render(bank, np.array([trajs[s][f] for s in range(n_streams)]), np.arange(n_streams) + 100 * f, offset=f * n_streams,
                      batch=n_streams)
    frames = bank.frames.reshape(n_frames, n_streams, cam.height, cam.width)
    dets = []
    for f in range(n_frames):
        row = []
        for s in range(n_streams):
            d = synth.detections(trajs[s][f], cam, rngs[s])
            if (f + 2 * s) % 7 == 6:
                d = d[:1]
            row.append(d)
        dets.append([torch.as_tensor(a, device=ctxvga.tdev) for a in pack_detections(row)])

    single = BatchedPoseDetector(ctxvga, n_streams, cam.width, cam.height, synth.object_points())
    ref_pose = torch.zeros((n_frames, n_streams, 6), dtype=torch.float64, device=ctxvga.tdev)
    ref_acc = []
    for f in range(n_frames):
        single.frames.copy_(frames[f])
        out = single.step_frames() if from_pixels else single.step(*dets[f])
        ref_pose[f].copy_(out["pose"])
        ref_acc.append(out["accepted"].clone())

    ctxs = [AgtContext(0, cam.mtx, None) for _ in range(groups)]
    for c in ctxs:
        c.set_synthetic_model()
    sg = StreamGroups(ctxs, n_streams, cam.width, cam.height, synth.object_points())
    assert [sl.stop - sl.start for sl in sg.slices] == [3, 4, 3]
    got_pose = torch.zeros_like(ref_pose)
    for rep in range(2):                                  # the second pass replays the captured graphs from a reset state
        sg.reset()
        sg.load(frames[0])
        sg.fork()
        accs = torch.zeros((n_frames, n_streams), dtype=torch.int32, device=ctxvga.tdev)
        for f in range(n_frames):
            nxt = frames[f + 1] if f + 1 < n_frames else None
            if from_pixels:
                sg.step(next_frames=nxt, pose_out=got_pose[f], accepted_out=accs[f])
            else:
                sg.step(*dets[f], next_frames=nxt, pose_out=got_pose[f], accepted_out=accs[f])
        sg.join()
        torch.cuda.synchronize()
        for f in range(n_frames):
            a = accs[f]
            assert torch.equal(a.cpu() != 0, ref_acc[f].reshape(-1).cpu() != 0), (rep, f)
            ok = (a != 0).cpu().numpy()
            assert np.array_equal(got_pose[f].cpu().numpy()[ok], ref_pose[f].cpu().numpy()[ok]), (rep, f)
        assert sum(int(x.sum()) for x in ref_acc) >= n_frames * n_streams * 0.8
    for c in ctxs:
        c.close()
