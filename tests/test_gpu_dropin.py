"""GPU tests of the reference-facing surface: the OpenCV-shaped host API (cv_compat, through the
``*_host`` C-ABI entry points), the drop-in ``PoseDetector`` and the committed golden vectors."""
import logging
import sys
from pathlib import Path

import numpy as np
import pytest

from accurate_aprilgroup_tracking_b200 import synth
from oracle import make_golden
from tests import util

pytestmark = pytest.mark.gpu
GOLDEN = make_golden.GOLDEN


@pytest.fixture(scope="module")
def host(lib_built):
    from accurate_aprilgroup_tracking_b200 import cv_compat
    h = cv_compat.default_context()
    s, tg, n, c = synth.surface_model()
    h.set_model(s, tg, n, c, synth.model_pitch())
    return h


def _logger():
    lg = logging.getLogger("agt-test")
    lg.handlers[:] = [logging.NullHandler()]
    lg.propagate = False
    return lg


class _Det:
    def __init__(self, tag_id, corners, margin=100.0):
        self.tag_id, self.corners, self.decision_margin = int(tag_id), np.asarray(corners, dtype=np.float64), margin
        self.center = self.corners.mean(axis=0)


@pytest.fixture()
def detector_factory(tmp_path, lib_built):
    from accurate_aprilgroup_tracking_b200.aprilgroup_pose_estimation import PoseDetector
    synth.write_april_group_json(tmp_path)

    def make(mtx, **kw):
        cls = type("PD", (PoseDetector,), {"DIRPATH": str(tmp_path / "aprilgroup_tracking" / "aprilgroup_pose_estimation")})
        return cls(_logger(), mtx, None, True, **kw)
    return make


# ------------------------------------------------------------------------------------------
def test_solvepnp_shapes_and_inplace_guess_semantics(host):
    import cv2
    cam = synth.CAMERA_VGA
    rng = np.random.default_rng(3)
    pose = synth.trajectory(11, 1)[0]
    dets = synth.detections(pose, cam, rng)
    obj = np.concatenate([synth.object_points()[4 * t:4 * t + 4] for t, _ in dets]).astype(np.float32)
    img = np.concatenate([c for _, c in dets]).astype(np.float32)
    ok, r, t = host.solvePnP(obj, img, cam.mtx, None, flags=0)
    okc, rc, tc = cv2.solvePnP(obj, img, cam.mtx, None, flags=cv2.SOLVEPNP_ITERATIVE)
    assert ok and r.shape == (3, 1) and t.shape == (3, 1) and r.dtype == np.float64
    util.assert_pose_close(np.concatenate([r.ravel(), t.ravel()]), np.concatenate([rc.ravel(), tc.ravel()]))
    # guess path: result is written into the caller's arrays, which are returned; float32 tvec stays float32
    g_r = (pose[:3] + 0.02).reshape(3, 1).copy()
    g_t = (pose[3:] + 0.002).astype(np.float32).reshape(3, 1)
    ok2, r2, t2 = host.solvePnP(obj, img, cam.mtx, None, g_r, g_t, True, flags=0)
    assert r2 is g_r and t2 is g_t and t2.dtype == np.float32
    util.assert_pose_close(np.concatenate([r2.ravel(), t2.ravel().astype(np.float64)]), np.concatenate([rc.ravel(), tc.ravel()]))
    with pytest.raises(ValueError):
        host.solvePnP(obj, img[:-1], cam.mtx, None)
    with pytest.raises(ValueError):
        host.solvePnP(obj, img, cam.mtx, None, flags=1)
    pts, _ = host.projectPoints(obj, r, t, cam.mtx, None)
    ref, _ = cv2.projectPoints(obj, rc, tc, cam.mtx, None)
    assert pts.shape == ref.shape and pts.dtype == ref.dtype and np.abs(pts - ref).max() < 1e-2


def test_lk_and_pyramid_golden(host):
    g = np.load(GOLDEN / "lk_pair.npz")
    nxt, st, err = host.calcOpticalFlowPyrLK(g["prev"], g["next"], g["pts"].reshape(-1, 1, 2), None)
    assert nxt.shape == (len(g["pts"]), 1, 2) and st.shape == (len(g["pts"]), 1) and st.dtype == np.uint8
    assert np.array_equal(st.ravel(), g["status"])
    m = g["status"] == 1
    assert np.abs(nxt.reshape(-1, 2)[m] - g["next_pts"][m]).max() <= util.FLOW_TOL
    assert np.abs(err.ravel()[m] - g["err"][m]).max() <= 0.05
    lv = host.pyramid(g["prev"], 4)
    for l in (1, 2, 3):
        assert np.array_equal(lv[l], g[f"level{l}"])
    assert np.array_equal(host.Scharr(g["prev"]), g["scharr0"])
    assert np.array_equal(host.Scharr(lv[2]), g["scharr2"])
    max_level, pyr = host.buildOpticalFlowPyramid(g["prev"], (21, 21), 3, True)
    assert max_level == 3 and len(pyr) == 8 and pyr[1].shape == g["prev"].shape + (2,)
    with pytest.raises(ValueError):
        host.calcOpticalFlowPyrLK(g["prev"], g["next"], g["pts"], None, winSize=(15, 15))


def test_dense_refinement_golden(host):
    g = np.load(GOLDEN / "dpr_case.npz")
    for i in range(len(g["frames"])):
        ok, r, t, cost, evals = host.refine_pose(g["frames"][i], g["init"][i][:3], g["init"][i][3:], g["mtx"])
        want = g["result"][i]
        assert ok
        util.assert_pose_close(np.concatenate([r.ravel(), t.ravel()]), want[:6], f"case {i}")
        assert abs(cost - want[6]) <= 1e-3 * want[6]
        assert abs(evals - int(want[8])) <= 2


def test_refine_host_chunked_equals_device_path(host, ctxvga):
    cam = synth.CAMERA_VGA
    rng = np.random.default_rng(77)
    n = 40
    truth = np.array([synth.random_pose(rng) for _ in range(n)])
    pyr = ctxvga.alloc_pyramid(n, cam.width, cam.height, 4)
    ctxvga.render(pyr, truth, np.arange(n) + 900)
    ctxvga.build_pyramid(pyr)
    init = truth + np.concatenate([rng.normal(0, 0.008, (n, 3)), rng.normal(0, 0.0004, (n, 3))], axis=1)
    dev = ctxvga.refine(pyr, init.reshape(n, 1, 6), 1)
    frames = pyr.frames.cpu().numpy()
    out = host.refine_poses(frames, init, cam.mtx)
    assert np.array_equal(out["pose"].reshape(n, 6), dev["pose"].cpu().numpy().reshape(n, 6))     # same kernels, same bits
    assert np.array_equal(out["evals"].ravel(), dev["evals"].cpu().numpy().ravel())
    # multi-hypothesis through the host entry point
    init3 = np.repeat(init[:6, None, :], 3, axis=1)
    init3[:, 1, :3] += 0.01
    out3 = host.refine_poses(frames[:6], init3, cam.mtx, n_hyp=3)
    score = 2.0 * out3["cost"].astype(np.float64) / np.maximum(out3["n_valid"], 1)
    # winner = lowest index within 1e-4 of the best score (oracle/dpr_oracle.py:refine_multi)
    want = [int(np.nonzero(s <= s.min() * (1 + 1e-4))[0][0]) for s in score]
    assert np.array_equal(out3["best"], want)


def test_refine_host_roi_upload_is_exact(host, ctxvga):
    """ROI-only upload (default) returns the same bits as uploading whole frames, moves far fewer bytes,
    and frames whose refinement leaves the predicted ROI are transparently redone from the full frame."""
    cam = synth.CAMERA_VGA
    rng = np.random.default_rng(99)
    n = 24
    truth = np.array([synth.random_pose(rng) for _ in range(n)])
    pyr = ctxvga.alloc_pyramid(n, cam.width, cam.height, 1)
    ctxvga.render(pyr, truth, np.arange(n) + 700)
    frames = pyr.frames.cpu().numpy()
    init = truth + np.concatenate([rng.normal(0, 0.008, (n, 3)), rng.normal(0, 0.0004, (n, 3))], axis=1)
    init[3, 3] += 0.012            # 12 mm off: the LM run starts (and wanders) outside the predicted ROI
    init[7, :3] += 0.25
    host.set_roi_upload(False)
    full = host.refine_poses(frames, init, cam.mtx)
    full_bytes = host.last_h2d_bytes()
    host.set_roi_upload(True)
    # poison the device staging buffers so stale pixels outside the ROI cannot help by accident
    host.refine_poses(np.full_like(frames, 255), init, cam.mtx)
    roi = host.refine_poses(frames, init, cam.mtx)
    roi_bytes = host.last_h2d_bytes()
    for key in ("pose", "cost", "n_valid", "evals", "status"):
        assert np.array_equal(roi[key], full[key]), key
    assert roi_bytes < 0.6 * full_bytes, (roi_bytes, full_bytes)
    # pinned host frames take the gather-kernel path (one launch per chunk instead of one 2-D copy per frame)
    from accurate_aprilgroup_tracking_b200 import cv_compat
    pinned = cv_compat.pinned_frames(frames.shape)              # an ordinary numpy array, in page-locked memory
    assert isinstance(pinned, np.ndarray) and pinned.dtype == np.uint8
    pinned[...] = frames
    host.refine_poses(np.full_like(frames, 255), init, cam.mtx)
    l0 = host.launch_count()
    pin = host.refine_poses(pinned, init, cam.mtx)
    assert host.launch_count() - l0 >= 2          # gather + refinement with its fused pyrDown (a batch this small runs as clusters,
                                                  # which derive their setup record themselves) (+ redo pass)
    for key in ("pose", "cost", "n_valid", "evals", "status"):
        assert np.array_equal(pin[key], full[key]), key
    assert host.last_h2d_bytes() == roi_bytes


def test_pose_detector_matches_reference_golden_sequence(detector_factory):
    """The drop-in class over the sequence the UNMODIFIED reference produced the golden states for:
    accept / reset decisions identical, poses within tolerance, aliasing quirks included."""
    g = np.load(GOLDEN / "ape_sequence.npz")
    det = detector_factory(g["mtx"])
    assert np.abs(det.all_objpts - g["all_objpts"]).max() < 1e-8
    for f in range(g["ids"].shape[0]):
        dets = [_Det(t, c) for t, c in make_golden.unpack_detections(g["ids"], g["corners"], f)]
        det.img = None
        img_list, obj_list, ids = det._lists_from_detections(dets)
        assert ids == [d.tag_id for d in dets]                       # tag ids / corner indexing bit-exact
        det._estimate_pose(img_list, obj_list)
        for key, attr in (("prev", det.prev_transform), ("guess", det.extrinsic_guess)):
            want = g[key][f]
            assert (attr[0] is None) == bool(np.isnan(want[0])), (f, key)
            if attr[0] is not None:
                util.assert_pose_close(np.concatenate([attr[0].ravel(), attr[1].ravel()]), want, f"{key} frame {f}")
        assert len(det.rot_velocities) == g["n_vel"][f]
    assert det.extrinsic_guess[1].dtype == np.float32                # predicted guess keeps the float32 tvec


def test_pose_detector_full_pipeline_with_lk_and_dense_refinement(detector_factory, ctxvga):
    """APE -> (LK when < 2 tags) -> dense refinement through the reference's per-frame entry point."""
    stub_dir = str(Path(__file__).resolve().parent.parent / "oracle" / "apriltag_stub")
    sys.path.insert(0, stub_dir)
    try:
        import apriltag as stub
        cam = synth.CAMERA_VGA
        det = detector_factory(cam.mtx, use_lk=True, use_dense_refine=True)
        assert det.options is not None
        n = 12
        traj = synth.trajectory(4321, n)
        rng = np.random.default_rng(4321)
        pyr = ctxvga.alloc_pyramid(n, cam.width, cam.height, 1)
        ctxvga.render(pyr, traj, np.arange(n) + 50)
        frames = pyr.frames.cpu().numpy()
        errs_t = []
        for f in range(n):
            dets = synth.detections(traj[f], cam, rng)
            margins = [100.0] * len(dets)
            if f in (6, 7):
                dets = dets[:1]                                       # detector loses all but one tag -> LK takes over
                margins = margins[:1]
            stub.push_detections(stub.Detection(t, c, m) for (t, c), m in zip(dets, margins))
            det._detect_and_get_pose(np.repeat(frames[f][:, :, None], 3, axis=2))
            assert det.prev_transform[0] is not None, f
            got = np.concatenate([det.prev_transform[0].ravel(), det.prev_transform[1].ravel().astype(np.float64)])
            dr, dt = util.pose_diff(got, traj[f])
            errs_t.append(dt)
            assert dr < 0.02 and dt < 2e-3, (f, dr, dt)
            if f in (6, 7):
                assert det.extrinsic_guess[0] is not None             # tracking survived the detection loss
                assert len(det._frame_corners) >= 2
        print("pipeline translation errors (mm)", np.round(np.array(errs_t) * 1e3, 3))
    finally:
        sys.path.remove(stub_dir)
        sys.modules.pop("apriltag", None)


def test_full_pipeline_matches_pipeline_oracle_config1(detector_factory, ctxvga):
    """BASELINE config 1 shape (640x480 synthetic sequence, APE + LK + dense refinement): the drop-in detector
    against the CPU composition of the three stage oracles, frame by frame, including detector dropouts."""
    from oracle import ape_oracle, pipeline_oracle
    cam = synth.CAMERA_VGA
    n = 300                                        # BASELINE config 1: the whole 300-frame sequence
    traj = synth.trajectory(1000, n)
    rng = np.random.default_rng(1000)
    pyr = ctxvga.alloc_pyramid(n, cam.width, cam.height, 1)
    ctxvga.render(pyr, traj, np.arange(n) + 1000)
    frames = pyr.frames.cpu().numpy()
    det = detector_factory(cam.mtx, use_lk=True, use_dense_refine=True)
    po = pipeline_oracle.PipelineOracle(ape_oracle.group_from_json(synth.april_group_dict()), cam.mtx, util.dpr_model())
    tracked = 0
    for f in range(n):
        dets = synth.detections(traj[f], cam, rng)
        if f in (12, 13, 27) or f % 37 == 36:
            dets = dets[:1]
        if f in (20, 150, 151):
            dets = []
        po.frame(frames[f], dets)
        det.img = None
        det._set_gray(frames[f])
        lists = det._lists_from_detections([_Det(t, c) for t, c in dets])
        before = len(lists[0])
        if len(lists[0]) < 2:
            lists = det._track_lost_tags(*lists)
        assert len(lists[0]) - before == po.tracked, f          # LK inlier set (all-four-corners rule) identical
        tracked += po.tracked
        det._estimate_pose(lists[0], lists[1])
        assert bool(det._prev_corners) == po.last_accepted, f
        assert (det.extrinsic_guess[0] is None) == (po.guess[0] is None), f
        if po.prev[0] is not None:
            got = np.concatenate([det.prev_transform[0].ravel(), det.prev_transform[1].ravel().astype(np.float64)])
            want = np.concatenate([po.prev[0].ravel(), po.prev[1].ravel().astype(np.float64)])
            util.assert_pose_close(got, want, f"frame {f}")
    assert tracked >= 10


def test_distorted_camera_and_arbitrary_tag_ids_through_process_frame(tmp_path, lib_built, ctxvga):
    """The reference's real flow: a calibrated camera with five distortion coefficients, frames through process_frame
    (cv.undistort + crop, detect_pose.py:611-619) and then _detect_and_get_pose (detect_pose.py:576-609), with LK and dense
    refinement on, and an april_group.json whose ids are neither contiguous nor in order.  Against the CPU composition of
    the stage oracles given the same frames (cv2.undistort, cv2.cvtColor) and detections.  Dense refinement runs on the
    undistorted pixels with the new camera matrix moved by the crop offset - it is never skipped."""
    import json
    import cv2
    from accurate_aprilgroup_tracking_b200.aprilgroup_pose_estimation import PoseDetector
    from oracle import ape_oracle, pipeline_oracle
    ids = [17, 4, 230, 9, 41, 3, 88, 12, 150, 7, 66, 21]                # id of the tag at position k of the JSON
    group = {"tags": {str(ids[k]): v for k, v in enumerate(synth.april_group_dict()["tags"].values())}}
    d = tmp_path / "aprilgroup_tracking" / "aprilgroup_pose_estimation"
    d.mkdir(parents=True)
    (d / "april_group.json").write_text(json.dumps(group))
    cam = synth.CAMERA_VGA
    dist = np.array([[-0.012, 0.004, 0.0003, -0.0002, 0.001]])
    w, h = cam.width, cam.height
    new_mtx, roi = cv2.getOptimalNewCameraMatrix(cam.mtx, dist, (w, h), 1, (w, h))
    pin = new_mtx.copy(); pin[0, 2] -= roi[0]; pin[1, 2] -= roi[1]
    pin_cam = synth.Camera(roi[2], roi[3], pin[0, 0], pin[1, 1], pin[0, 2], pin[1, 2])
    # raw (distorted) frames: the pinhole render seen through the lens model
    gx, gy = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    und = cv2.undistortPoints(np.stack([gx, gy], axis=-1).reshape(-1, 1, 2), cam.mtx, dist, P=cam.mtx).reshape(h, w, 2)
    n = 14
    traj = synth.trajectory(777, n)
    pyr = ctxvga.alloc_pyramid(n, w, h, 1)
    ctxvga.render(pyr, traj, np.arange(n) + 300)
    pinhole = pyr.frames.cpu().numpy()
    stub_dir = str(Path(__file__).resolve().parent.parent / "oracle" / "apriltag_stub")
    sys.path.insert(0, stub_dir)
    try:
        import apriltag as stub
        cls = type("PD", (PoseDetector,), {"DIRPATH": str(d)})
        det = cls(_logger(), cam.mtx, dist, True, use_lk=True, use_dense_refine=True)
        assert list(det.extrinsics) == ids
        po = pipeline_oracle.PipelineOracle(ape_oracle.group_from_json(group), cam.mtx, util.dpr_model(), dist=dist)
        rng = np.random.default_rng(777)
        refined = 0
        for f in range(n):
            raw = cv2.remap(pinhole[f], und[..., 0], und[..., 1], cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=128)
            raw = np.repeat(raw[:, :, None], 3, axis=2)
            raw[..., 0] = np.clip(raw[..., 0].astype(np.int32) + 3, 0, 255)      # channels differ: the gray weights matter
            # what the reference computes on the host
            want_frame = cv2.undistort(raw, cam.mtx, dist, None, new_mtx)[roi[1]:roi[1] + roi[3], roi[0]:roi[0] + roi[2]]
            want_gray = cv2.cvtColor(want_frame, cv2.COLOR_BGR2GRAY)
            frame = det.process_frame(raw)
            assert np.array_equal(frame, want_frame), f                          # cv.undistort + crop, bit-exact, 3 channels
            # detections on the undistorted frame, reported under the JSON's ids
            dets = [(ids[k], c) for k, c in synth.detections(traj[f], pin_cam, rng)]
            if f in (6, 7):
                dets = dets[:1]                                                  # LK has to carry the other tags
            stub.push_detections(stub.Detection(t, c, 100.0) for t, c in dets)
            det._detect_and_get_pose(frame)
            assert np.array_equal(det._gray, want_gray), f                       # gray from the same device pass, bit-exact
            po.frame(want_gray, dets, refine_mtx=pin)
            assert bool(det._prev_corners) == po.last_accepted, f
            assert (det.extrinsic_guess[0] is None) == (po.guess[0] is None), f
            if po.last_accepted:
                refined += 1
                got = np.concatenate([det.prev_transform[0].ravel(), det.prev_transform[1].ravel().astype(np.float64)])
                want = np.concatenate([po.prev[0].ravel(), po.prev[1].ravel().astype(np.float64)])
                util.assert_pose_close(got, want, f"frame {f}")
                dr, dt = util.pose_diff(got, traj[f])
                assert dr < 0.02 and dt < 2e-3, (f, dr, dt)                      # and it is the right pose
        assert refined >= 10
        # a distorted frame that did not go through process_frame cannot be refined: loud, not skipped
        stub.push_detections(stub.Detection(t, c, 100.0) for t, c in dets)
        with pytest.raises(ValueError, match="process_frame"):
            det._detect_and_get_pose(raw)
        with pytest.raises(KeyError):
            det._lists_from_detections([_Det(5, np.zeros((4, 2)))])               # id 5 is not in this group
    finally:
        sys.path.remove(stub_dir)
        sys.modules.pop("apriltag", None)


def test_undistort_ingest_golden(host):
    # the reference's frame ingest (cv2 calls of detect_pose.py:167-177, 602) as committed golden vectors
    g = np.load(GOLDEN / "undistort_case.npz")
    for k in range(2):
        x, y, w, h = (int(v) for v in g["roi"][k])
        got = host.undistort_gray(g["frame"], g["mtx"], g["dist"][k], g["new_mtx"][k], (x, y, w, h))
        assert got.shape == (h, w) and np.array_equal(got.ravel(), g[f"gray{k}"])

